"""Long-form decode in time chunks with recompute halos (BASELINE config 5).

The decoder's receptive field is +-11.5 latent frames for the shipped hyper-parameters,
so a chunk decoded with a 12-frame halo on each side reproduces the unchunked result on its interior.
Halos are CLIPPED at the utterance ends (never zero-padded): the reference zero-pads every layer's input
at the true boundary, which only the true first / last chunk may see.  Interior chunks share one shape
and are decoded as ONE batch, so a 60 s utterance becomes a handful of batched launches.
"""
import torch


def chunk_plan(frames, chunk_frames, halo):
    """[(lo, hi, keep_lo, keep_hi)]: decode z[lo:hi], keep output frames [keep_lo, keep_hi) of that piece."""
    plan = []
    s = 0
    while s < frames:
        e = min(frames, s + chunk_frames)
        lo, hi = max(0, s - halo), min(frames, e + halo)
        plan.append((lo, hi, s - lo, s - lo + (e - s)))
        s = e
    return plan


def decode_chunked(decode_fn, z, g=None, chunk_frames=512, halo=12, hop=256):
    """z: [B, C, T].  decode_fn(z, g) -> [B', 1, T'*hop].  Returns [B, 1, T*hop]."""
    B, C, T = z.shape
    if T <= chunk_frames + 2 * halo:
        return decode_fn(z, g)
    plan = chunk_plan(T, chunk_frames, halo)
    out = torch.empty((B, 1, T * hop), dtype=torch.float32, device=z.device)
    groups = {}
    for idx, (lo, hi, klo, khi) in enumerate(plan):
        groups.setdefault((hi - lo, klo, khi), []).append(idx)
    for (length, klo, khi), idxs in groups.items():
        zs = torch.cat([z[:, :, plan[i][0]:plan[i][1]] for i in idxs], dim=0)      # [len(idxs)*B, C, length]
        gs = None if g is None else g.repeat(len(idxs), 1, 1)
        y = decode_fn(zs, gs)
        for n, i in enumerate(idxs):
            s = plan[i][0] + klo
            out[:, :, s * hop:(s + khi - klo) * hop] = y[n * B:(n + 1) * B, :, klo * hop:khi * hop]
    return out

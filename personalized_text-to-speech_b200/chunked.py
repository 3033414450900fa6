"""Long-form decode in time chunks with recompute halos (BASELINE config 5).

A chunk decoded with ``halo`` extra latent frames on each side reproduces the unchunked result on its interior when
``halo`` covers the decoder's receptive field.  ``receptive_halo`` computes that field exactly from the
hyper-parameters by back-propagating the needed input interval through conv_post, the ResBlock stacks, the
ConvTranspose1d stages and conv_pre (models.py:270-289): **13 frames** on either side for the shipped configuration
(conv_post 3 samples; per stage the widest branch, k = 11: 5 * (1+3+5 + 3*1) = 60 samples; the transposed convs divide
by their stride; conv_pre 3 frames), verified bit for bit in tests/test_generator_host.py (a latent frame 13 away
changes a frame's samples, one 14 away does not).  With >= 13 the chunked result equals the unchunked one up to
summation order (1e-16 in fp64); 12 leaves ~2e-10 (the taps that far out carry almost no weight), which is why the
earlier "+-11.5 frames, use 12" estimate passed every fp32 / bf16 test.  ``decode_chunked`` defaults to the exact value.
Halos are CLIPPED at the utterance ends (never zero-padded): the reference zero-pads every layer's input
at the true boundary, which only the true first / last chunk may see.  ``decode_chunked`` uses windows of one length
(``chunk_plan_uniform``: the first and last windows are shifted inward rather than clipped), so a 60 s utterance is ONE
batched decode.
"""
import torch


def receptive_field(resblock, resblock_kernel_sizes, resblock_dilation_sizes, upsample_rates, upsample_kernel_sizes):
    """(left, right): latent frames before / after frame f that the output samples of frame f depend on (exact, by
    interval back-propagation; (13, 13) for the shipped configuration)."""
    left = right = 3                                 # conv_post: k = 7, models.py:264
    for s, k in zip(reversed(list(upsample_rates)), reversed(list(upsample_kernel_sizes))):
        widest = 0
        for rk, dil in zip(resblock_kernel_sizes, resblock_dilation_sizes):
            hk = (int(rk) - 1) // 2
            if str(resblock) == "1":                 # modules.py:210-223: c1 (dilation d) then c2 (dilation 1), 3 times
                widest = max(widest, hk * (sum(int(d) for d in dil[:3]) + 3))
            else:                                    # modules.py:246-252
                widest = max(widest, hk * sum(int(d) for d in dil))
        s, k = int(s), int(k)
        p = (k - s) // 2                             # ConvTranspose1d: o = s*i - p + j, j in [0, k)
        left = (left + widest + k - 1 - p) // s      # i_min = ceil((o_min + p - (k-1)) / s)
        right = (s - 1 + right + widest + p) // s    # i_max = floor((o_max + p) / s)
    return left + 3, right + 3                       # conv_pre: k = 7, models.py:249


def receptive_halo(resblock, resblock_kernel_sizes, resblock_dilation_sizes, upsample_rates, upsample_kernel_sizes):
    """Halo (frames on each side of a chunk) that covers the receptive field: max(left, right)."""
    return max(receptive_field(resblock, resblock_kernel_sizes, resblock_dilation_sizes, upsample_rates,
                               upsample_kernel_sizes))


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


def chunk_plan_uniform(frames, chunk_frames, halo):
    """The same cover with windows of ONE length W = chunk_frames + 2*halo: [(lo, lo + W, keep_lo, keep_hi)].

    Interior chunks are chunk_plan's.  The first and last windows are shifted inward instead of being clipped -- they
    carry more than `halo` frames of context on their inner side and end exactly at the utterance boundary, where the
    reference's own zero padding applies -- so every chunk of an utterance has the same shape and the whole utterance is
    ONE batched decode (before: the interior chunks as one batch plus two single-chunk decodes of their own lengths, 4.06 ms
    for 60 s against 3.37 ms unchunked on a B200; BASELINE config 5).  Needs frames >= W."""
    W = chunk_frames + 2 * halo
    assert frames >= W
    plan = []
    s = 0
    while s < frames:
        e = min(frames, s + chunk_frames)
        lo = min(max(0, s - halo), frames - W)
        plan.append((lo, lo + W, s - lo, e - lo))
        s = e
    return plan


def decode_chunked(decode_fn, z, g=None, chunk_frames=512, halo=None, hop=256):
    """z: [B, C, T].  decode_fn(z, g) -> [B', 1, T'*hop].  Returns [B, 1, T*hop].

    halo=None: the exact receptive field of ``decode_fn`` when it is a ``Generator`` (``receptive_halo()``), else the
    shipped configuration's 13 frames."""
    if halo is None:
        rh = getattr(decode_fn, "receptive_halo", None)
        halo = rh() if callable(rh) else 13
    B, C, T = z.shape
    if T <= chunk_frames + 2 * halo:
        return decode_fn(z, g)
    plan = chunk_plan_uniform(T, chunk_frames, halo)   # every window has the same length: one batched decode
    zs = torch.cat([z[:, :, lo:hi] for (lo, hi, _, _) in plan], dim=0)             # [len(plan)*B, C, W]
    gs = None if g is None else g.repeat(len(plan), 1, 1)
    y = decode_fn(zs, gs)
    out = torch.empty((B, 1, T * hop), dtype=y.dtype, device=y.device)
    for n, (lo, hi, klo, khi) in enumerate(plan):
        s = lo + klo
        out[:, :, s * hop:(s + khi - klo) * hop] = y[n * B:(n + 1) * B, :, klo * hop:khi * hop]
    return out

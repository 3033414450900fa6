"""Utterance sharding across the GPUs of one box (BASELINE config 4).

Utterances are independent on this path (no cross-batch op in models.py:270-289), so rank r of N decodes
utterances [r*B/N, (r+1)*B/N) with replicated decoder weights and NO collective inside the decoder.
The only exchange is the final gather of fp32 waveforms, and only when the caller wants every rank
(or rank 0) to hold the whole batch.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous balanced split: the first n % world ranks get one extra utterance."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def decode_sharded(decode_fn, z, g=None, group=None, gather="all"):
    """Decode this rank's slice of a replicated batch and (optionally) gather the waveforms.

    decode_fn(z_slice, g_slice) -> [b, 1, L] is ``Generator.forward`` on a GPU rank (tests inject the CPU
    oracle to exercise the sharding logic under gloo).  ``gather``: "all" (every rank gets [B,1,L]),
    "none" (each rank keeps its own slice; returns (slice, (lo, hi))).
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = z.shape[0]
    lo, hi = shard_range(B, rank, world)
    if hi > lo:
        mine = decode_fn(z[lo:hi], None if g is None else g[lo:hi])
    else:
        mine = None
    if gather == "none" or world == 1:
        return mine if world == 1 else (mine, (lo, hi))
    # equal-size fast path: one all_gather_into_tensor of [B/N, 1, L]
    sizes = [shard_range(B, r, world) for r in range(world)]
    counts = [b - a for a, b in sizes]
    L = torch.tensor([0 if mine is None else mine.shape[-1]], device=z.device, dtype=torch.int64)
    dist.all_reduce(L, op=dist.ReduceOp.MAX, group=group)
    L = int(L.item())
    if mine is None:
        mine = z.new_zeros((0, 1, L), dtype=torch.float32)
    if len(set(counts)) == 1:
        out = torch.empty((B, 1, L), dtype=mine.dtype, device=mine.device)
        dist.all_gather_into_tensor(out, mine.contiguous(), group=group)
        return out
    pad = max(counts)
    buf = mine.new_zeros((pad, 1, L))
    buf[: mine.shape[0]] = mine
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)

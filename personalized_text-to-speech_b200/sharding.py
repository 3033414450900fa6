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


class PeerGather:
    """The final waveform gather as copy-engine PUSHES over NVLink peer memory: no kernel, no SM.

    Why not ``all_gather_into_tensor`` when the gather overlaps the next batch's decode: the decoder's kernels are persistent
    launches of one CTA per SM with a static tile assignment, and an NCCL kernel that spins for its peer holds a few SMs
    while it waits -- every decode launch beside it then needs two rounds (bench.py at N = 2: 18.4 ms per step instead of
    8.5; serialising gather and decode instead costs a cross-rank rendezvous per step, 9.6 ms).  Here every rank owns a
    symmetric buffer ``[slots][world][...]`` (``torch.distributed._symmetric_memory``: the same allocation mapped into
    every rank of the box) and copies its own waveforms into row ``rank`` of EVERY rank's buffer with plain device-to-device
    copies on a side stream: 14 MB per peer at NVLink speed (0.04 ms), issued by the copy engines while the SMs decode the
    next batch.  ``finish()`` is the only synchronisation: a barrier behind the last push, after which ``full(slot)`` holds
    the whole batch on every rank.  Needs CUDA peer access between the ranks' GPUs (one NVSwitch box): the
    constructor raises where the rendezvous is not possible, callers then fall back to ``decode_sharded``'s NCCL gather.
    """

    def __init__(self, shape_per_rank, dtype, device, group=None, slots=2):
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.buf = symm_mem.empty((slots, self.world) + tuple(shape_per_rank), dtype=dtype, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, group)
        self.peers = [self.hdl.get_buffer(r, self.buf.shape, dtype) for r in range(self.world)]
        self.stream = torch.cuda.Stream(device=device)
        self.slots = slots

    def push(self, y, slot, after=None):
        """Copy this rank's waveforms ``y`` into row ``rank`` of slot ``slot`` on every rank, behind the work already
        enqueued on ``after`` (default: the current stream).  Returns immediately."""
        after = after or torch.cuda.current_stream(y.device)
        ev = torch.cuda.Event()
        ev.record(after)
        self.stream.wait_event(ev)
        y.record_stream(self.stream)
        with torch.cuda.stream(self.stream):
            for r in range(self.world):
                self.peers[r][slot % self.slots, self.rank].copy_(y, non_blocking=True)

    def finish(self, stream=None):
        """Barrier behind every rank's pushes; afterwards ``stream`` (default: current) may read ``full(slot)``."""
        stream = stream or torch.cuda.current_stream(self.buf.device)
        with torch.cuda.stream(self.stream):
            self.hdl.barrier()
        stream.wait_stream(self.stream)

    def full(self, slot):
        """[world * b, ...] view of a slot: the whole batch in rank order."""
        t = self.buf[slot % self.slots]
        return t.reshape((t.shape[0] * t.shape[1],) + tuple(t.shape[2:]))

"""2-GPU probe: does torch's symmetric memory rendezvous work on this box, and what does a copy-engine push of one rank's
waveforms (14 MB) into a peer's buffer cost while the pusher's SMs are busy?  torchrun --nproc-per-node 2 tools/peer_gather_probe.py"""
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm_mem

n = 16 * 862 * 256
try:
    buf = symm_mem.empty((2, world, n), dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
    peers = [hdl.get_buffer(r, buf.shape, buf.dtype) for r in range(world)]
except Exception as e:
    print("rank %d: symmetric memory unavailable: %r" % (rank, e))
    dist.destroy_process_group()
    sys.exit(0)
y = torch.full((n,), float(rank + 1), device=dev)
cs = torch.cuda.Stream()
torch.cuda.synchronize()
dist.barrier()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(cs):
    ev0.record()
    for it in range(10):
        for r in range(world):
            peers[r][it % 2, rank].copy_(y, non_blocking=True)
    ev1.record()
torch.cuda.synchronize()
hdl.barrier()
torch.cuda.synchronize()
ok = all(float(buf[1, r, 0]) == r + 1 and float(buf[1, r, -1]) == r + 1 for r in range(world))
print("rank %d: push of %.1f MB to %d ranks: %.3f ms per step; contents ok: %s" % (rank, n * 4 / 1e6, world, ev0.elapsed_time(ev1) / 10, ok))
dist.destroy_process_group()

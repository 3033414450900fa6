"""CPU, world_size 2 over gloo: utterance sharding + final waveform gather reproduce the unsharded decode."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, B, q):
    sys.path.insert(0, ROOT)
    import oracle
    import vitsdec
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    hp = oracle.TINY
    sd = oracle.synth_state_dict(hp, 7, gain=2.0)
    rs = np.random.RandomState(3)
    z = torch.from_numpy(rs.standard_normal((B, hp.initial_channel, 6)).astype(np.float32))
    g = torch.from_numpy(rs.standard_normal((B, hp.gin_channels, 1)).astype(np.float32))
    calls = []

    def fn(zz, gg):
        calls.append(zz.shape[0])
        return torch.from_numpy(oracle.generator_forward_np(hp, sd, zz.numpy(), gg.numpy(), dtype=np.float32))

    out = vitsdec.decode_sharded(fn, z, g, gather="all")
    full = fn(z, g)
    lo, hi = vitsdec.shard_range(B, rank, world)
    q.put((rank, float((out - full).abs().max()), calls[0], hi - lo))
    dist.destroy_process_group()


def _run(B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(res)


def test_even_split_gathers_bit_identical():
    for rank, err, decoded, mine in _run(4):
        assert err == 0.0 and decoded == mine == 2


def test_ragged_split_gathers_bit_identical():
    res = _run(3)
    assert [r[3] for r in res] == [2, 1]
    assert all(r[1] == 0.0 for r in res)

"""Generate the committed golden vectors by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

For each case it
  1. synthesises a state_dict with oracle.synth_state_dict (numpy RandomState, so the GPU
     box regenerates bit-identical weights from the seed),
  2. loads it (strict) into /root/reference/models_infer.Generator -- byte-identical to
     models.Generator but importable without building monotonic_align,
  3. runs Generator.forward(z, g) in fp32 on CPU,
  4. stores inputs, output, a weight checksum and a few intermediates in an .npz.

It also checks, right here, that both oracle restatements reproduce the reference
(numpy fp64 within fp32 round-off; torch-functional bit-identical), and stores per-op
vectors for conv1d / conv_transpose1d / weight_norm taken from torch itself.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

import models_infer  # noqa: E402  (the reference)
import oracle  # noqa: E402
from oracle.generator_torch import generator_forward_torch, to_torch_state_dict  # noqa: E402

from tests.golden.cases import CASES, weight_checksum  # noqa: E402


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    only = sys.argv[1:]   # optional name filters: regenerate just the matching cases (the .npz bytes carry timestamps)
    for name, hp, seed, B, T, use_g in CASES:
        if only and not any(o in name for o in only):
            continue
        sd = oracle.synth_state_dict(hp, seed, gain=2.0)
        args, kw = hp.ctor_args()
        G = models_infer.Generator(*args, **kw).eval()
        ref_keys = list(G.state_dict().keys())
        assert ref_keys == [k for k, _ in oracle.state_dict_keys(hp)], "state_dict key order mismatch"
        G.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
        rs = np.random.RandomState(seed + 1000)
        z = rs.standard_normal((B, hp.initial_channel, T)).astype(np.float32)
        g = rs.standard_normal((B, hp.gin_channels, 1)).astype(np.float32) if use_g else None
        with torch.no_grad():
            y = G(torch.from_numpy(z), g=None if g is None else torch.from_numpy(g)).numpy()
        # the restatements, checked against the real thing
        taps = {}
        y_np = oracle.generator_forward_np(hp, sd, z, g, dtype=np.float64, taps=taps)
        y_t = generator_forward_torch(hp, to_torch_state_dict(sd), torch.from_numpy(z),
                                      None if g is None else torch.from_numpy(g)).numpy()
        err_np = np.abs(y_np - y).max()
        err_t = np.abs(y_t - y).max()
        print("%-18s out %s  max|y| %.4f rms %.4f  |np64-ref| %.2e  |torchF-ref| %.2e" % (
            name, y.shape, np.abs(y).max(), np.sqrt((y ** 2).mean()), err_np, err_t))
        assert err_np < 5e-6 and err_t < 5e-6, (err_np, err_t)
        out = dict(z=z, y=y, seed=np.int64(seed), gain=np.float64(2.0), wsum=weight_checksum(sd),
                   mrf0_mean=np.float64(taps["mrf.0"].mean()), mrf0_std=np.float64(taps["mrf.0"].std()))
        if g is not None:
            out["g"] = g
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)

    if only:
        return
    # per-op vectors straight from torch (the third-party arithmetic the reference composes)
    rs = np.random.RandomState(7)
    x = rs.standard_normal((2, 6, 17)).astype(np.float32)
    w = rs.standard_normal((4, 6, 5)).astype(np.float32)
    b = rs.standard_normal((4,)).astype(np.float32)
    wt = rs.standard_normal((6, 3, 8)).astype(np.float32)
    bt = rs.standard_normal((3,)).astype(np.float32)
    gg = rs.uniform(0.5, 1.5, (6, 1, 1)).astype(np.float32)
    ops = dict(
        x=x, w=w, b=b, wt=wt, bt=bt, gg=gg,
        conv_d3=torch.nn.functional.conv1d(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b),
                                           dilation=3, padding=6).numpy(),
        convt_s4=torch.nn.functional.conv_transpose1d(torch.from_numpy(x), torch.from_numpy(wt),
                                                      torch.from_numpy(bt), stride=4, padding=2).numpy(),
        wn=torch._weight_norm(torch.from_numpy(wt), torch.from_numpy(gg), 0).numpy(),
        lrelu=torch.nn.functional.leaky_relu(torch.from_numpy(x), 0.1).numpy(),
    )
    np.savez_compressed(os.path.join(HERE, "ops.npz"), **ops)
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()

"""Golden vectors of the flow: the UNMODIFIED reference ResidualCouplingBlock run in the build container.

    python tests/golden/make_golden_flow.py

For each case: seeded numpy weights (oracle.flow_torch.synth_flow_state_dict) loaded strictly into
/root/reference/models_infer.ResidualCouplingBlock, forward(x, x_mask, g, reverse) in fp32 on the CPU; inputs and output
stored in tests/golden/<name>.npz.  The restatement (oracle/flow_torch.py) is checked against the reference right here.
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
from oracle.flow_torch import flow_forward_torch, flow_state_dict_keys, synth_flow_state_dict  # noqa: E402
from tests.golden.flow_cases import FLOW_CASES  # noqa: E402


def main():
    torch.set_num_threads(8)
    for name, hp, seed, B, T, lengths, reverse in FLOW_CASES:
        sd = synth_flow_state_dict(hp, seed)
        args, kw = hp.ctor_args()
        F = models_infer.ResidualCouplingBlock(*args, **kw).eval()
        assert list(F.state_dict().keys()) == [k for k, _ in flow_state_dict_keys(hp)], "state_dict key order mismatch"
        tsd = {k: torch.from_numpy(v) for k, v in sd.items()}
        F.load_state_dict(tsd, strict=True)
        rs = np.random.RandomState(seed + 1000)
        x = rs.standard_normal((B, hp.channels, T)).astype(np.float32)
        g = rs.standard_normal((B, hp.gin_channels, 1)).astype(np.float32) if hp.gin_channels else None
        lens = np.array(lengths if lengths is not None else [T] * B, dtype=np.int64)
        mask = (np.arange(T)[None, :] < lens[:, None]).astype(np.float32)[:, None, :]
        with torch.no_grad():
            y = F(torch.from_numpy(x), torch.from_numpy(mask), g=None if g is None else torch.from_numpy(g),
                  reverse=reverse).numpy()
            y2 = flow_forward_torch(hp, tsd, torch.from_numpy(x), torch.from_numpy(mask),
                                    None if g is None else torch.from_numpy(g), reverse=reverse).numpy()
        err = np.abs(y - y2).max()
        # how far the couplings move x (a flow whose post layers are zero would only permute channels)
        moved = np.abs(y - (x[:, ::-1] if hp.n_flows % 2 else x)).max()
        print("%-26s out %s  max|y| %.3f  |restatement-ref| %.2e  coupling shift %.3f" % (name, y.shape, np.abs(y).max(),
                                                                                      err, moved))
        assert err < 1e-5 and moved > 0.05
        out = dict(x=x, y=y, lens=lens, seed=np.int64(seed), reverse=np.int64(reverse))
        if g is not None:
            out["g"] = g
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


if __name__ == "__main__":
    main()

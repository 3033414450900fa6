"""CPU: the flow restatement (oracle/flow_torch.py) against the committed outputs of the unmodified reference
ResidualCouplingBlock (tests/golden/flow_*.npz, made by tests/golden/make_golden_flow.py), and the host side of the
drop-in module (parameter tree, argument checks)."""
import importlib
import os

import numpy as np
import pytest
import torch

from oracle.flow_torch import (FLOW_FINETUNE_SPEAKER, FLOW_TINY, flow_forward_torch, flow_state_dict_keys,
                               synth_flow_state_dict)
from tests.golden.flow_cases import FLOW_CASES

flowmod = importlib.import_module("personalized_text-to-speech_b200.flow")


def _mask(lens, T):
    return (torch.arange(T)[None, :] < torch.as_tensor(lens)[:, None]).float()[:, None, :]


@pytest.mark.parametrize("case", FLOW_CASES, ids=[c[0] for c in FLOW_CASES])
def test_flow_restatement_matches_reference(golden_dir, case):
    name, hp, seed, B, T, lengths, reverse = case
    gold = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    sd = {k: torch.from_numpy(v) for k, v in synth_flow_state_dict(hp, seed).items()}
    g = torch.from_numpy(gold["g"]) if "g" in gold else None
    y = flow_forward_torch(hp, sd, torch.from_numpy(gold["x"]), _mask(gold["lens"], T), g, reverse=bool(gold["reverse"]))
    assert y.shape == (B, hp.channels, T)
    assert float((y - torch.from_numpy(gold["y"])).abs().max()) < 5e-6


def test_flow_reverse_inverts_forward():
    """Size-independent property of a coupling flow: reverse(forward(x)) == x on the valid frames."""
    hp = FLOW_TINY
    sd = {k: torch.from_numpy(v) for k, v in synth_flow_state_dict(hp, 3).items()}
    x = torch.randn(2, hp.channels, 13)
    g = torch.randn(2, hp.gin_channels, 1)
    m = _mask([13, 6], 13)
    z = flow_forward_torch(hp, sd, x, m, g, reverse=False)
    xr = flow_forward_torch(hp, sd, z, m, g, reverse=True)
    assert float(((xr - x) * m).abs().max()) < 1e-4


def test_dropin_module_has_the_reference_parameter_tree():
    for hp in (FLOW_FINETUNE_SPEAKER, FLOW_TINY):
        args, kw = hp.ctor_args()
        F = flowmod.ResidualCouplingBlock(*args, **kw)
        want = flow_state_dict_keys(hp)
        got = [(k, tuple(v.shape)) for k, v in F.state_dict().items()]
        assert got == [(k, tuple(s)) for k, s in want]
        F.load_state_dict({k: torch.from_numpy(v) for k, v in synth_flow_state_dict(hp, 1).items()}, strict=True)


def test_dropin_module_refuses_cpu_tensors_and_grad():
    hp = FLOW_TINY
    args, kw = hp.ctor_args()
    F = flowmod.ResidualCouplingBlock(*args, **kw).eval()
    x = torch.randn(1, hp.channels, 4)
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU path"):
        F(x, torch.ones(1, 1, 4))
    with pytest.raises(RuntimeError, match="expected x of shape"):
        F(torch.randn(1, hp.channels + 2, 4), torch.ones(1, 1, 4))

"""CPU: the oracle restatements against the committed reference outputs (tests/golden/*.npz,
produced by tests/golden/make_golden.py from the unmodified /root/reference Generator)."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import generator_np as gnp
from oracle.generator_torch import generator_forward_torch, to_torch_state_dict
from tests.golden.cases import CASES, weight_checksum


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name + ".npz")))


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_generator_np_matches_reference(golden_dir, case):
    name, hp, seed, B, T, use_g = case
    gold = _load(golden_dir, name)
    sd = oracle.synth_state_dict(hp, seed, gain=float(gold["gain"]))
    # the weights the GPU box regenerates from the seed are the ones the reference ran with
    assert np.isclose(weight_checksum(sd), gold["wsum"], rtol=0, atol=1e-9)
    if name.startswith("full") and T > 8:
        dtype = np.float32  # keep the CPU suite fast; fp64 is checked on the short cases
        tol = 2e-5
    else:
        dtype = np.float64
        tol = 5e-6
    y = oracle.generator_forward_np(hp, sd, gold["z"], gold.get("g"), dtype=dtype)
    assert y.shape == gold["y"].shape == (B, 1, T * hp.hop)
    assert np.abs(y - gold["y"]).max() < tol


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_generator_torch_matches_reference(golden_dir, case):
    name, hp, seed, B, T, use_g = case
    gold = _load(golden_dir, name)
    sd = to_torch_state_dict(oracle.synth_state_dict(hp, seed, gain=float(gold["gain"])))
    g = gold.get("g")
    y = generator_forward_torch(hp, sd, torch.from_numpy(gold["z"]), None if g is None else torch.from_numpy(g))
    assert np.abs(y.numpy() - gold["y"]).max() < 2e-6


def test_folded_and_weight_norm_forms_agree(golden_dir):
    name, hp, seed, B, T, use_g = CASES[0]
    gold = _load(golden_dir, name)
    sd = oracle.synth_state_dict(hp, seed, gain=2.0)
    folded = oracle.weights.fold_state_dict(sd)
    assert len(folded) == len(sd) - sum(k.endswith("weight_g") for k in sd)
    y = oracle.generator_forward_np(hp, folded, gold["z"], gold.get("g"))
    assert np.abs(y - gold["y"]).max() < 5e-6


def test_state_dict_key_counts():
    # 233 tensors with weight norm, 157 after remove_weight_norm
    assert len(oracle.state_dict_keys(oracle.FINETUNE_SPEAKER, True)) == 233
    assert len(oracle.weights.fold_state_dict(oracle.synth_state_dict(oracle.TINY, 0))) == \
        len(oracle.state_dict_keys(oracle.TINY, False))
    sd = oracle.synth_state_dict(oracle.FINETUNE_SPEAKER, 0)
    assert sum(v.size for v in sd.values()) == 14468608  # SURVEY.md section 8a (weight-norm form)


def test_ops_match_torch(golden_dir):
    o = _load(golden_dir, "ops")
    assert np.abs(gnp.conv1d(o["x"], o["w"], o["b"], dilation=3, padding=6) - o["conv_d3"]).max() < 1e-5
    assert np.abs(gnp.conv_transpose1d(o["x"], o["wt"], o["bt"], stride=4, padding=2) - o["convt_s4"]).max() < 1e-5
    assert np.abs(oracle.fold_weight_norm(o["wt"], o["gg"]) - o["wn"]).max() < 1e-6
    assert np.array_equal(gnp.leaky_relu(o["x"], np.float32(0.1)), o["lrelu"])


def test_weight_norm_axis_is_dim0_for_transposed_conv():
    # ConvTranspose1d weight is [C_in, C_out, k]; weight_g is [C_in,1,1] (models.py:254)
    rs = np.random.RandomState(0)
    v = rs.standard_normal((6, 3, 8)).astype(np.float32)
    g = rs.uniform(0.5, 1.5, (6, 1, 1)).astype(np.float32)
    w = oracle.fold_weight_norm(v, g)
    assert np.allclose(np.sqrt((w.astype(np.float64) ** 2).sum(axis=(1, 2))), g[:, 0, 0], rtol=1e-6)


def test_edge_lengths():
    hp = oracle.TINY
    sd = oracle.synth_state_dict(hp, 3)
    for T in (1, 2, 7):
        z = np.random.RandomState(T).standard_normal((1, hp.initial_channel, T))
        y = oracle.generator_forward_np(hp, sd, z, None)
        assert y.shape == (1, 1, T * hp.hop) and np.isfinite(y).all() and np.abs(y).max() <= 1.0


def test_chunked_decode_halo_property():
    """a 12-frame halo reproduces the unchunked output (receptive field
    +-11.5 latent frames for the shipped config); shown here on the tiny config's own field."""
    hp = oracle.TINY
    sd = oracle.synth_state_dict(hp, 5, gain=2.0)
    T, chunk, halo = 40, 10, 12
    rs = np.random.RandomState(9)
    z = rs.standard_normal((1, hp.initial_channel, T))
    full = oracle.generator_forward_np(hp, sd, z, None)
    parts = []
    for s in range(0, T, chunk):
        lo, hi = max(0, s - halo), min(T, s + chunk + halo)
        y = oracle.generator_forward_np(hp, sd, z[:, :, lo:hi], None)
        parts.append(y[:, :, (s - lo) * hp.hop:(s - lo + min(chunk, T - s)) * hp.hop])
    assert np.abs(np.concatenate(parts, axis=2) - full).max() < 1e-9

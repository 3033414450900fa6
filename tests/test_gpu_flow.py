"""GPU: parity of the native flow (vitsdec_flow_apply through ResidualCouplingBlock) against the committed outputs of
the unmodified reference and against the fp32 restatement at larger sizes, plus the flow -> decoder chain that
SynthesizerTrn.infer runs (models.py:521-522).

Stated tolerance (bf16 conv operands, fp32 accumulation, fp32 latent): SNR >= 35 dB, max-abs error <= 3 % of the peak.
"""
import os

import numpy as np
import pytest
import torch

import oracle
import vitsdec
from oracle.flow_torch import (FLOW_FINETUNE_SPEAKER, FLOW_TINY, flow_forward_torch, synth_flow_state_dict)
from oracle.generator_torch import generator_forward_torch, to_torch_state_dict
from tests.golden.flow_cases import FLOW_CASES

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def snr_db(ref, got):
    ref, got = ref.double(), got.double()
    return 10 * np.log10(float((ref ** 2).sum()) / max(float(((ref - got) ** 2).sum()), 1e-300))


def check(ref, got, snr_min=35.0, frac=0.03):
    assert got.shape == ref.shape and torch.isfinite(got).all()
    s = snr_db(ref, got)
    m = float((ref - got).abs().max()) / float(ref.abs().max())
    assert s >= snr_min and m <= frac, "SNR %.1f dB, max-abs %.3f of peak" % (s, m)
    return s, m


def mask_of(lens, T, device="cpu"):
    return (torch.arange(T, device=device)[None, :] < torch.as_tensor(lens, device=device)[:, None]).float()[:, None, :]


def build(hp, seed):
    args, kw = hp.ctor_args()
    F = vitsdec.ResidualCouplingBlock(*args, **kw)
    sd = synth_flow_state_dict(hp, seed)
    F.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return F.to(DEV).eval(), {k: torch.from_numpy(v) for k, v in sd.items()}


@pytest.mark.parametrize("case", FLOW_CASES, ids=[c[0] for c in FLOW_CASES])
def test_golden_reference_outputs(golden_dir, case):
    name, hp, seed, B, T, lengths, reverse = case
    gold = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    F, _ = build(hp, seed)
    x = torch.from_numpy(gold["x"]).to(DEV)
    g = torch.from_numpy(gold["g"]).to(DEV) if "g" in gold else None
    with torch.no_grad():
        y = F(x, mask_of(gold["lens"], T, DEV), g=g, reverse=bool(gold["reverse"]))
    check(torch.from_numpy(gold["y"]), y.cpu())


@pytest.mark.parametrize("B,T,lens,reverse", [(16, 862, None, True), (3, 301, (301, 17, 150), True),
                                              (2, 1000, (1000, 999), False), (1, 1, None, True)])
def test_full_config_vs_fp32_restatement(B, T, lens, reverse):
    hp = FLOW_FINETUNE_SPEAKER
    F, sd = build(hp, 60 + B)
    rs = np.random.RandomState(B * 7 + T)
    x = torch.from_numpy(rs.standard_normal((B, hp.channels, T)).astype(np.float32))
    g = torch.from_numpy(rs.standard_normal((B, hp.gin_channels, 1)).astype(np.float32))
    m = mask_of(lens if lens is not None else [T] * B, T)
    with torch.no_grad():
        y = F(x.to(DEV), m.to(DEV), g=g.to(DEV), reverse=reverse).cpu()
    ref = flow_forward_torch(hp, sd, x, m, g, reverse=reverse)
    check(ref, y)
    # frames past an utterance's length: x0 passes through, x1 is zeroed, exactly like `* x_mask` does
    assert torch.equal(y * (1 - m) != 0, ref * (1 - m) != 0)


@pytest.mark.parametrize("n_layers", [1, 3, 5, 6])
def test_layer_counts_around_the_skip_sum_segment_limit(n_layers):
    """Up to 4 WN layers the skip sum of a coupling is ONE 4-segment launch accumulated in TMEM (flow.cu
    FlowCoupling::skip); deeper WNs keep the per-layer split epilogue with the fp32 skip tensor.  Both against the fp32
    restatement, both directions, ragged mask, dilation_rate 2."""
    from oracle.flow_torch import FlowHParams
    hp = FlowHParams(64, 64, 3, 2, n_layers, 2, 32)
    F, sd = build(hp, 90 + n_layers)
    B, T = 3, 200
    rs = np.random.RandomState(n_layers)
    x = torch.from_numpy(rs.standard_normal((B, hp.channels, T)).astype(np.float32))
    g = torch.from_numpy(rs.standard_normal((B, hp.gin_channels, 1)).astype(np.float32))
    m = mask_of((200, 57, 133), T)
    for reverse in (False, True):
        with torch.no_grad():
            y = F(x.to(DEV), m.to(DEV), g=g.to(DEV), reverse=reverse).cpu()
        check(flow_forward_torch(hp, sd, x, m, g, reverse=reverse), y)


def test_reverse_inverts_forward_on_device():
    hp = FLOW_FINETUNE_SPEAKER
    F, _ = build(hp, 70)
    x = torch.randn(2, hp.channels, 400, device=DEV)
    g = torch.randn(2, hp.gin_channels, 1, device=DEV)
    m = mask_of([400, 123], 400, DEV)
    with torch.no_grad():
        z = F(x, m, g=g, reverse=False)
        xr = F(z, m, g=g, reverse=True)
    assert snr_db((x * m).cpu(), (xr * m).cpu()) > 35.0


def test_no_speaker_conditioning_and_half_input():
    hp = FLOW_TINY
    F, sd = build(hp, 71)
    x = torch.randn(2, hp.channels, 50)
    m = mask_of([50, 31], 50)
    with torch.no_grad():
        y = F(x.to(DEV), m.to(DEV), g=None, reverse=True).cpu()
        yh = F(x.to(DEV).half(), m.to(DEV), g=None, reverse=True)
    check(flow_forward_torch(hp, sd, x, m, None, reverse=True), y)
    assert yh.dtype == torch.float16
    # any binary mask, not only a sequence mask: `* x_mask` is applied row by row like the reference does
    holes = m.clone()
    holes[0, 0, 3] = 0
    holes[1, 0, 10:14] = 0
    with torch.no_grad():
        yh2 = F(x.to(DEV), holes.to(DEV), g=None, reverse=True).cpu()
        y_none = F(x.to(DEV), None, g=None, reverse=True).cpu()
    check(flow_forward_torch(hp, sd, x, holes, None, reverse=True), yh2)
    check(flow_forward_torch(hp, sd, x, torch.ones_like(m), None, reverse=True), y_none)


def test_flow_then_decoder_matches_the_reference_chain():
    """z = flow(z_p, y_mask, g, reverse=True); o = dec(z * y_mask, g) -- models.py:521-522, both stages native."""
    fhp, ghp = FLOW_FINETUNE_SPEAKER, oracle.FINETUNE_SPEAKER
    F, fsd = build(fhp, 72)
    args, kw = ghp.ctor_args()
    G = vitsdec.Generator(*args, **kw)
    gsd = oracle.synth_state_dict(ghp, 73, gain=2.0)
    G.load_state_dict({k: torch.from_numpy(v) for k, v in gsd.items()})
    G = G.to(DEV).eval()
    B, T = 2, 120
    rs = np.random.RandomState(5)
    z_p = torch.from_numpy(rs.standard_normal((B, fhp.channels, T)).astype(np.float32))
    g = torch.from_numpy(rs.standard_normal((B, fhp.gin_channels, 1)).astype(np.float32))
    m = mask_of([T, 77], T)
    with torch.no_grad():
        z = F(z_p.to(DEV), m.to(DEV), g=g.to(DEV), reverse=True)
        o = G(z * m.to(DEV), g.to(DEV)).cpu()
    z_ref = flow_forward_torch(fhp, fsd, z_p, m, g, reverse=True)
    o_ref = generator_forward_torch(ghp, to_torch_state_dict(gsd), z_ref * m, g)
    check(z_ref, z.cpu())
    check(o_ref, o, snr_min=30.0, frac=0.05)   # two bf16 stages in series


@pytest.mark.parametrize("B,T,lens,reverse", [(3, 301, (301, 17, 150), True), (2, 1000, (1000, 999), False)])
def test_fp16_mode_vs_fp32_restatement_and_back(B, T, lens, reverse):
    """Option "fp16": fp16 conv operands / stored activations.  Stated tolerance for this mode: SNR >= 55 dB (the bf16
    mode's bound is 35 dB, measured 45-50)."""
    hp = FLOW_FINETUNE_SPEAKER
    F, sd = build(hp, 60 + B)
    rs = np.random.RandomState(B * 7 + T)
    x = torch.from_numpy(rs.standard_normal((B, hp.channels, T)).astype(np.float32))
    g = torch.from_numpy(rs.standard_normal((B, hp.gin_channels, 1)).astype(np.float32))
    m = mask_of(lens, T)
    ref = flow_forward_torch(hp, sd, x, m, g, reverse=reverse)
    with torch.no_grad():
        y_bf = F(x.to(DEV), m.to(DEV), g=g.to(DEV), reverse=reverse).cpu()
        F.set_option("fp16", 1)
        y_h = F(x.to(DEV), m.to(DEV), g=g.to(DEV), reverse=reverse).cpu()
        F.set_option("fp16", 0)
        y_bf2 = F(x.to(DEV), m.to(DEV), g=g.to(DEV), reverse=reverse).cpu()
    s_bf, _ = check(ref, y_bf)
    s_h, _ = check(ref, y_h, snr_min=55.0, frac=0.005)
    assert s_h > s_bf + 10.0
    assert torch.equal(y_bf, y_bf2)


def test_fp16_mode_flow_then_decoder_chain():
    fhp, ghp = FLOW_FINETUNE_SPEAKER, oracle.FINETUNE_SPEAKER
    F, fsd = build(fhp, 72)
    F.set_option("fp16", 1)
    args, kw = ghp.ctor_args()
    G = vitsdec.Generator(*args, **kw)
    gsd = oracle.synth_state_dict(ghp, 73, gain=2.0)
    G.load_state_dict({k: torch.from_numpy(v) for k, v in gsd.items()})
    G = G.to(DEV).eval()
    G.set_option("fp16", 1)
    B, T = 2, 120
    rs = np.random.RandomState(5)
    z_p = torch.from_numpy(rs.standard_normal((B, fhp.channels, T)).astype(np.float32))
    g = torch.from_numpy(rs.standard_normal((B, fhp.gin_channels, 1)).astype(np.float32))
    m = mask_of([T, 77], T)
    with torch.no_grad():
        z = F(z_p.to(DEV), m.to(DEV), g=g.to(DEV), reverse=True)
        o = G(z * m.to(DEV), g.to(DEV)).cpu()
    z_ref = flow_forward_torch(fhp, fsd, z_p, m, g, reverse=True)
    o_ref = generator_forward_torch(ghp, to_torch_state_dict(gsd), z_ref * m, g)
    check(z_ref, z.cpu(), snr_min=55.0, frac=0.005)
    check(o_ref, o, snr_min=45.0, frac=0.01)   # two fp16 stages in series (bf16: 30 dB)

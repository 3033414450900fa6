"""GPU: the drop-in claim at the level the reference's own scripts use it (north_star: "SynthesizerTrn.infer,
cmd_inference.py and VC_inference.py run unchanged").

The UNMODIFIED reference (baseline/_ref, installed byte for byte by baseline/install_ref.py) is imported on the B200 and
its ``models.SynthesizerTrn`` is built exactly as cmd_inference.py:92-99 builds it (configs/finetune_speaker.json,
random init, seed 1234), once stock and once after ``vitsdec.patch_reference()`` -- i.e. with ``models.Generator``
(models.py:447) resolved to the B200 decoder.  Both run

  * ``infer`` (models.py:499-523) on the same 50 synthetic symbol ids, sid 0, same RNG seed, the way
    cmd_inference.py:109-114 calls it,
  * ``voice_conversion`` (models.py:525-533),
  * a ``utils.save_checkpoint`` -> ``utils.load_checkpoint`` round trip (utils.py:148-193) into the patched model,

and the waveforms are compared: SNR >= 35 dB and max-abs <= 3 % of the peak (the decoder's stated bf16 tolerance).  The
latent the stock ``dec`` saw is captured with a forward-pre-hook and fed to the patched ``dec`` as well, so the decoder
comparison does not depend on the rest of the model being bit-reproducible.  With ``patch_reference(flow=True)`` the flow
(models.py:521) is native too: two bf16 stages in series, stated 30 dB / 5 % (as tests/test_gpu_flow.py).
"""
import json
import os

import numpy as np
import pytest
import torch

import vitsdec
from tests import refinstall

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(refinstall.ref_dir() is None, reason="baseline/_ref not installed "
                                                                      "(python baseline/install_ref.py)")]
DEV = "cuda:0"
N_SYMBOLS = 68   # len(symbols) of the trilingual cleaner set the shipped configs were trained with


def snr_db(ref, got):
    ref, got = ref.double(), got.double()
    return 10 * np.log10(float((ref ** 2).sum()) / max(float(((ref - got) ** 2).sum()), 1e-300))


def check(ref, got, snr_min=35.0, frac=0.03):
    assert got.shape == ref.shape and got.dtype == ref.dtype
    assert torch.isfinite(got).all()
    s = snr_db(ref, got)
    m = float((ref - got).abs().max()) / float(ref.abs().max())
    assert s >= snr_min and m <= frac, "SNR %.1f dB, max-abs %.3f of peak" % (s, m)
    return s, m


def _cfg(name="finetune_speaker.json"):
    return json.load(open(os.path.join(refinstall.ref_dir(), "configs", name)))


def _build(models, cfg, seed=1234):
    """cmd_inference.py:93-99."""
    torch.manual_seed(seed)
    net = models.SynthesizerTrn(N_SYMBOLS, cfg["data"]["filter_length"] // 2 + 1,
                                cfg["train"]["segment_size"] // cfg["data"]["hop_length"],
                                n_speakers=cfg["data"]["n_speakers"], **cfg["model"]).to(DEV)
    net.eval()
    return net


def _randomise(net, seed=7):
    """Random init leaves weight_g == ||weight_v|| (a fold that ignored g would pass) and zero-initialises the flow's
    ``post`` convs (modules.py:320-321: the couplings would be identities): perturb both, identically for every build."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    with torch.no_grad():
        for name, p in sorted(net.named_parameters()):
            if name.startswith(("dec.", "flow.")) and name.endswith("weight_g"):
                p.mul_(torch.empty(p.shape).uniform_(0.5, 1.5, generator=gen).to(p.device))
            if name.startswith("flow.") and (name.endswith("post.weight") or name.endswith("post.bias")):
                p.copy_(torch.empty(p.shape).uniform_(-0.05, 0.05, generator=gen).to(p.device))


@pytest.fixture(scope="module")
def ref_models():
    models = refinstall.import_reference("models")
    # the stock model is the fp32 reference here: no TF32 convs / matmuls (torch's cuDNN default would round the stock
    # decoder's operands to 10 mantissa bits), deterministic algorithms so both runs draw the same durations
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    yield models
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic = saved
    vitsdec.unpatch_reference()
    refinstall.forget_reference()


def _pair(models, flow=False, cfg_name="finetune_speaker.json"):
    """(stock model, patched model with the stock model's weights)."""
    cfg = _cfg(cfg_name)
    stock = _build(models, cfg)
    _randomise(stock)
    assert type(stock.dec).__module__ == "models"
    vitsdec.patch_reference(("models",), flow=flow)
    try:
        ours = _build(models, cfg)
    finally:
        vitsdec.unpatch_reference()
    assert isinstance(ours.dec, vitsdec.Generator)
    assert isinstance(ours.flow, vitsdec.ResidualCouplingBlock) == flow
    ours.load_state_dict(stock.state_dict(), strict=True)
    return stock, ours


def _infer(net, x, xl, sid, seed=99):
    """cmd_inference.py:113-114."""
    torch.manual_seed(seed)
    with torch.no_grad():
        return net.infer(x, xl, sid=sid, noise_scale=.667, noise_scale_w=0.6, length_scale=1.0)


def _text(n=50, batch=1, seed=3):
    rs = np.random.RandomState(seed)
    x = torch.from_numpy(rs.randint(1, N_SYMBOLS, size=(batch, n))).long().to(DEV)
    xl = torch.full((batch,), n, dtype=torch.long, device=DEV)
    return x, xl


@pytest.mark.parametrize("flow", [False, True], ids=["dec", "flow+dec"])
def test_infer_runs_unchanged_and_matches_the_stock_model(ref_models, flow):
    stock, ours = _pair(ref_models, flow=flow)
    x, xl = _text()
    sid = torch.zeros(1, dtype=torch.long, device=DEV)
    seen = {}
    h = stock.dec.register_forward_pre_hook(lambda m, a, kw: seen.update(z=a[0].clone(), g=kw["g"].clone()),
                                            with_kwargs=True)
    o_ref, attn_ref, mask_ref, _ = _infer(stock, x, xl, sid)
    h.remove()
    o, attn, mask, (z, z_p, m_p, logs_p) = _infer(ours, x, xl, sid)
    # same output contract: [B, 1, 256 * T'], fp32, on the device (cmd_inference.py:114 indexes [0][0,0])
    assert o.dtype == torch.float32 and o.is_cuda and o.dim() == 3 and o.shape[1] == 1
    assert o.shape == o_ref.shape and o.shape[2] == 256 * mask.shape[2]
    assert torch.equal(mask, mask_ref)
    tol = dict(snr_min=30.0, frac=0.05) if flow else {}
    check(o_ref.cpu(), o.cpu(), **tol)
    audio = (o, attn, mask)[0][0, 0].data.cpu().float().numpy()     # cmd_inference.py:114 indexes the returned tuple
    assert audio.shape == (o.shape[2],) and np.isfinite(audio).all()
    # the decoder alone on exactly the latent the stock decoder saw
    with torch.no_grad():
        o2 = ours.dec(seen["z"], g=seen["g"])
    check(o_ref.cpu(), o2.cpu())


def test_infer_batch_of_three_with_max_len_and_ragged_lengths(ref_models):
    """B = 3, ragged text lengths (padded frames are decoded like any other, models.py:522) and the max_len slice the
    training-time eval uses (finetune_speaker_v2.py:331)."""
    stock, ours = _pair(ref_models)
    x, _ = _text(n=40, batch=3, seed=5)
    xl = torch.tensor([40, 17, 29], dtype=torch.long, device=DEV)
    sid = torch.tensor([0, 5, 998], dtype=torch.long, device=DEV)
    torch.manual_seed(11)
    with torch.no_grad():
        o_ref = stock.infer(x, xl, sid=sid, noise_scale=.667, noise_scale_w=0.8, length_scale=1.0, max_len=32)[0]
    torch.manual_seed(11)
    with torch.no_grad():
        o = ours.infer(x, xl, sid=sid, noise_scale=.667, noise_scale_w=0.8, length_scale=1.0, max_len=32)[0]
    assert o.shape == o_ref.shape and o.shape[0] == 3 and o.shape[2] <= 32 * 256
    check(o_ref.cpu(), o.cpu())


def test_voice_conversion_matches_the_stock_model(ref_models):
    """models.py:525-533: enc_q -> flow -> flow(reverse) -> dec."""
    cfg = _cfg()
    for flow, tol in ((False, {}), (True, dict(snr_min=30.0, frac=0.05))):
        stock, ours = _pair(ref_models, flow=flow)
        rs = np.random.RandomState(8)
        spec_ch = cfg["data"]["filter_length"] // 2 + 1
        y = torch.from_numpy(np.abs(rs.standard_normal((2, spec_ch, 60))).astype(np.float32)).to(DEV)
        yl = torch.tensor([60, 41], dtype=torch.long, device=DEV)
        src = torch.tensor([3, 4], dtype=torch.long, device=DEV)
        tgt = torch.tensor([10, 0], dtype=torch.long, device=DEV)
        torch.manual_seed(21)
        with torch.no_grad():
            o_ref, mask_ref, _ = stock.voice_conversion(y, yl, src, tgt)
        torch.manual_seed(21)
        with torch.no_grad():
            o, mask, _ = ours.voice_conversion(y, yl, src, tgt)
        assert torch.equal(mask, mask_ref) and o.shape == (2, 1, 60 * 256)
        check(o_ref.cpu(), o.cpu(), **tol)


def test_checkpoint_round_trip_through_the_reference_utils(ref_models, tmp_path):
    """utils.save_checkpoint (utils.py:183-193) of the STOCK model -> utils.load_checkpoint (utils.py:148-180) into a
    freshly built PATCHED model: load_checkpoint walks the model's own keys and silently keeps the init value of any
    key missing from the file, so this only reproduces the stock waveform if every dec.* / flow.* key matches."""
    utils = refinstall.import_reference("utils")
    cfg = _cfg()
    stock = _build(ref_models, cfg)
    _randomise(stock)
    path = str(tmp_path / "G_test.pth")
    utils.save_checkpoint(stock, None, 2e-4, 7, path)
    vitsdec.patch_reference(("models",), flow=True)
    try:
        ours = _build(ref_models, cfg, seed=4321)     # different init: everything must come from the file
    finally:
        vitsdec.unpatch_reference()
    _, _, lr, it = utils.load_checkpoint(path, ours, None)
    assert (lr, it) == (2e-4, 7)
    for k, v in stock.state_dict().items():
        assert torch.equal(v, ours.state_dict()[k]), k
    x, xl = _text()
    sid = torch.zeros(1, dtype=torch.long, device=DEV)
    o_ref = _infer(stock, x, xl, sid)[0]
    o = _infer(ours, x, xl, sid)[0]
    check(o_ref.cpu(), o.cpu(), snr_min=30.0, frac=0.05)
    # a second checkpoint loaded into the SAME live model must replace the folded weights (version fingerprint)
    other = _build(ref_models, cfg, seed=555)
    _randomise(other, seed=9)
    utils.save_checkpoint(other, None, 1e-4, 8, path)
    utils.load_checkpoint(path, ours, None)
    o_ref2 = _infer(other, x, xl, sid)[0]
    o2 = _infer(ours, x, xl, sid)[0]
    check(o_ref2.cpu(), o2.cpu(), snr_min=30.0, frac=0.05)


def test_uma_trilingual_config_builds_and_infers(ref_models):
    """BASELINE config 5's hyper-parameters (configs/uma_trilingual.json: same model block, n_speakers 999)."""
    stock, ours = _pair(ref_models, cfg_name="uma_trilingual.json")
    x, xl = _text(n=30, seed=9)
    sid = torch.tensor([123], dtype=torch.long, device=DEV)
    o_ref = _infer(stock, x, xl, sid)[0]
    o = _infer(ours, x, xl, sid)[0]
    check(o_ref.cpu(), o.cpu())

"""CPU: host-side mirror of the reference interface (state_dict contract, error behaviour, launcher, helpers)."""
import sys
import types

import numpy as np
import pytest
import torch

import oracle
import vitsdec
from vitsdec import Generator

chunked = __import__("importlib").import_module("personalized_text-to-speech_b200.chunked")


def make(hp=oracle.FINETUNE_SPEAKER):
    args, kw = hp.ctor_args()
    return Generator(*args, **kw)


def test_state_dict_keys_match_reference_tree():
    G = make()
    ref = oracle.state_dict_keys(oracle.FINETUNE_SPEAKER)
    sd = G.state_dict()
    assert list(sd.keys()) == [k for k, _ in ref]          # 233 keys, same order (utils.py:155-177 walks them)
    assert all(tuple(sd[k].shape) == tuple(s) for k, s in ref)
    assert len(sd) == 233


def test_strict_load_of_reference_state_dict_and_folded_form():
    hp = oracle.TINY
    G = make(hp)
    sd = oracle.synth_state_dict(hp, 3)
    G.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    for k, v in G.state_dict().items():
        assert np.array_equal(v.numpy(), sd[k])
    # the 157-key style (remove_weight_norm) form is accepted too and reproduces the same effective weights
    folded = oracle.weights.fold_state_dict(sd)
    G2 = make(hp)
    G2.load_state_dict({k: torch.from_numpy(v) for k, v in folded.items()}, strict=True)
    w = torch._weight_norm(G2.ups[0].weight_v, G2.ups[0].weight_g, 0).detach().numpy()
    assert np.allclose(w, folded["ups.0.weight"], rtol=1e-5, atol=1e-7)


def test_remove_weight_norm_changes_keys_like_reference(capsys):
    G = make(oracle.TINY)
    n_wn = len(G.state_dict())
    G.remove_weight_norm()
    assert "Removing weight norm" in capsys.readouterr().out      # models.py:292 prints this
    keys = list(G.state_dict().keys())
    assert len(keys) == len(oracle.state_dict_keys(oracle.TINY, weight_norm=False)) < n_wn
    assert not any(k.endswith("weight_g") for k in keys)


def test_resblock2_tree():
    G = make(oracle.TINY_RB2)
    assert [k for k, _ in oracle.state_dict_keys(oracle.TINY_RB2)] == list(G.state_dict().keys())
    assert not hasattr(G, "cond")                                   # gin_channels == 0, models.py:267


def test_forward_errors_are_loud_not_fallbacks():
    G = make(oracle.TINY).eval()
    with torch.no_grad():
        with pytest.raises(RuntimeError, match="no CPU path"):
            G(torch.zeros(1, 64, 4))
        with pytest.raises(RuntimeError, match="expected x of shape"):
            G(torch.zeros(1, 63, 4))


def test_patch_reference_replaces_generator_symbol():
    fake = types.ModuleType("models")
    fake.Generator = object
    sys.modules["models"] = fake
    try:
        done = vitsdec.patch_reference(("models",))
        assert done == ["models"] and fake.Generator is Generator
        vitsdec.unpatch_reference()
        assert fake.Generator is object
    finally:
        del sys.modules["models"]


def test_shard_range_partitions():
    for n in (0, 1, 7, 16, 256):
        for world in (1, 2, 3, 8):
            spans = [vitsdec.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_chunk_plan_covers_once_and_clips_halos():
    for T, c, h in ((5168, 512, 12), (100, 32, 12), (33, 32, 12), (40, 10, 3)):
        plan = chunked.chunk_plan(T, c, h)
        covered = []
        for lo, hi, klo, khi in plan:
            assert 0 <= lo < hi <= T and 0 <= klo < khi <= hi - lo
            covered += list(range(lo + klo, lo + khi))
            assert lo == max(0, lo + klo - h) and hi == min(T, lo + khi + h)
        assert covered == list(range(T))


def test_uniform_chunk_plan_covers_once_with_equal_windows():
    """decode_chunked's plan: windows of ONE length (one batched decode), every frame kept exactly once, at least `halo`
    frames of context on every side that is not the utterance boundary."""
    for (T, c, h) in ((5168, 512, 13), (100, 10, 12), (35, 10, 12), (1000, 8, 13), (539, 512, 13), (2 * 512 + 26, 512, 13)):
        plan = chunked.chunk_plan_uniform(T, c, h)
        kept = []
        for lo, hi, klo, khi in plan:
            assert 0 <= lo and hi <= T and hi - lo == c + 2 * h
            assert (klo >= h or lo == 0) and ((hi - lo) - khi >= h or hi == T)
            kept += list(range(lo + klo, lo + khi))
        assert kept == list(range(T))


def test_decode_chunked_with_oracle_decode_fn():
    hp = oracle.TINY
    sd = oracle.synth_state_dict(hp, 5, gain=2.0)
    z = torch.from_numpy(np.random.RandomState(2).standard_normal((1, hp.initial_channel, 50)).astype(np.float32))

    def fn(zz, gg):
        return torch.from_numpy(oracle.generator_forward_np(hp, sd, zz.numpy(), None, dtype=np.float64)).float()

    full = fn(z, None)
    got = vitsdec.decode_chunked(fn, z, None, chunk_frames=10, halo=12, hop=hp.hop)
    assert got.shape == full.shape and float((got - full).abs().max()) < 1e-6


def test_receptive_halo_is_exact_for_the_shipped_configuration():
    """the receptive field of the shipped decoder is 13 latent frames on either side (not "11.5, use 12"):
    with halo >= 13 the chunked decode equals the unchunked one up to fp64 summation order, 12 leaves ~2e-10; the
    computed value is the default of decode_chunked."""
    hp = oracle.FINETUNE_SPEAKER
    args, _ = hp.ctor_args()
    assert chunked.receptive_halo(args[1], args[2], args[3], args[4], args[6]) == 13
    assert chunked.receptive_halo("2", (3, 5, 7), ((1, 2), (2, 6), (3, 12)), (8, 8, 4), (16, 16, 8)) > 0
    G = vitsdec.Generator(*args, gin_channels=hp.gin_channels)
    assert G.receptive_halo() == 13
    from oracle.generator_torch import generator_forward_torch, to_torch_state_dict
    sd = to_torch_state_dict(oracle.synth_state_dict(hp, 3, gain=2.0), dtype=torch.float64)
    rs = np.random.RandomState(4)
    T = 48
    z = torch.from_numpy(rs.standard_normal((1, hp.initial_channel, T)))
    g = torch.from_numpy(rs.standard_normal((1, hp.gin_channels, 1)))

    def fn(zz, gg):
        return generator_forward_torch(hp, sd, zz, gg)

    full = fn(z, g)
    err = {}
    for halo in (12, 13, 14):
        got = vitsdec.decode_chunked(fn, z, g, chunk_frames=8, halo=halo, hop=hp.hop)
        err[halo] = float((got - full).abs().max())
    peak = float(full.abs().max())
    assert err[13] <= 1e-13 * peak and err[14] <= 1e-13 * peak, err     # summation-order noise only
    assert err[12] > 1e4 * err[13], err            # 12 frames do NOT cover the field (~2e-10: invisible in fp32)
    # the dependency itself, bit for bit: frame f0's samples see latent frames f0 - 13 .. f0 + 13 and nothing beyond
    assert chunked.receptive_field(args[1], args[2], args[3], args[4], args[6]) == (13, 13)
    f0 = 20
    for dist, depends in ((14, False), (13, True), (-14, False), (-13, True)):
        z2 = z.clone()
        z2[0, :, f0 + dist] += 1e3
        same = torch.equal(fn(z2, g)[0, 0, f0 * hp.hop:(f0 + 1) * hp.hop], full[0, 0, f0 * hp.hop:(f0 + 1) * hp.hop])
        assert same != depends, (dist, same)


def test_polyphase_transposed_conv_derivation():
    """The packing rule used by pack_convT_kernel (csrc/pack.cu): output sample s*i + r reads input rows i + off with
    kernel index j = r + p - s*off.  Restated in numpy and checked against the oracle's conv_transpose1d."""
    from oracle.generator_np import conv_transpose1d
    rs = np.random.RandomState(0)
    for (ci, co, k, s) in ((6, 4, 16, 8), (5, 3, 4, 2), (4, 2, 8, 4), (3, 2, 6, 2), (3, 2, 3, 1)):
        p = (k - s) // 2
        L = 11
        x = rs.standard_normal((1, ci, L))
        w = rs.standard_normal((ci, co, k))
        ref = conv_transpose1d(x, w, None, stride=s, padding=p)
        off_min = -((k - 1 - p) // s)
        off_max = (s - 1 + p) // s
        y = np.zeros((1, co, L * s))
        for off in range(off_min, off_max + 1):
            for r in range(s):
                j = r + p - s * off
                if 0 <= j < k:
                    for i in range(L):
                        if 0 <= i + off < L:
                            y[0, :, s * i + r] += w[:, :, j].T @ x[0, :, i + off]
        assert ref.shape == y.shape and np.abs(ref - y).max() < 1e-12


def _fold_weights(w, r):
    """numpy restatement of pack_conv_fold_kernel (csrc/pack.cu): W'[s][phi*C+co][psi*C+ci] = W[co][ci][j],
    j - (k-1)/2 = r*s + psi - phi."""
    co_n, ci_n, k = w.shape
    hk = (k - 1) // 2
    s_min, s_max = -((hk + r - 1) // r), (r - 1 + hk) // r
    out = np.zeros((s_max - s_min + 1, r * co_n, r * ci_n))
    for si, s in enumerate(range(s_min, s_max + 1)):
        for phi in range(r):
            for psi in range(r):
                j = r * s + psi - phi + hk
                if 0 <= j < k:
                    out[si, phi * co_n:(phi + 1) * co_n, psi * ci_n:(psi + 1) * ci_n] = w[:, :, j]
    return out, s_min


def test_time_folded_conv_derivation():
    """The time-folded form of decoder.cu fold_geom: a k-tap conv on [L][C] equals a conv over folded rows [L/r][r*C]
    with block-Toeplitz weights; a dilation-d conv equals the same folded conv on each sub-sequence t = d*q + rho
    (zero padding at both ends, ragged last row of a sub-sequence treated as zeros)."""
    from oracle.generator_np import conv1d
    rs = np.random.RandomState(1)
    for (C, k, r, d, L) in ((4, 3, 4, 1, 24), (3, 7, 2, 1, 18), (4, 11, 4, 1, 40), (4, 3, 4, 3, 29), (3, 7, 2, 5, 31),
                            (2, 11, 4, 2, 5)):
        x = rs.standard_normal((1, C, L))
        w = rs.standard_normal((C, C, k))
        ref = conv1d(x, w, None, dilation=d, padding=(k - 1) // 2 * d)[0].T        # [L][C]
        wf, s_min = _fold_weights(w, r)
        y = np.zeros((L, C))
        rows = -(-L // (d * r))
        for rho in range(d):
            # folded rows of sub-sequence rho: row n, phase phi <-> sample t = d*(r*n + phi) + rho  (zeros past the end)
            xf = np.zeros((rows, r * C))
            for n in range(rows):
                for phi in range(r):
                    t = d * (r * n + phi) + rho
                    if t < L:
                        xf[n, phi * C:(phi + 1) * C] = x[0, :, t]
            for n in range(rows):
                acc = np.zeros(r * C)
                for si in range(wf.shape[0]):
                    m = n + s_min + si
                    if 0 <= m < rows:
                        acc += wf[si] @ xf[m]
                for phi in range(r):
                    t = d * (r * n + phi) + rho
                    if t < L:
                        y[t] = acc[phi * C:(phi + 1) * C]
        assert np.abs(ref - y).max() < 1e-12, (C, k, r, d, L)


def test_options_are_remembered_until_a_native_handle_exists():
    """set_option before the first forward (no handle, no GPU) is recorded and replayed at vitsdec_create; changing
    the 16-bit storage format ("fp16") invalidates the folded weights so that the next forward re-folds them."""
    G = make(oracle.TINY)
    assert G._handle is None
    G._loaded_fingerprint = ("stale",)
    G.set_option("fp16", 1)
    assert G._options == {"fp16": 1} and G._loaded_fingerprint is None and G._handle is None
    G._loaded_fingerprint = ("kept",)
    G.set_option("fp16", 1)          # unchanged value: nothing to re-fold
    G.set_option("graph", 0)         # other options never touch the weights
    assert G._loaded_fingerprint == ("kept",) and G._options == {"fp16": 1, "graph": 0}
    G.set_option("fp16", 0)
    assert G._loaded_fingerprint is None
    F = vitsdec.ResidualCouplingBlock(8, 8, 5, 1, 2, n_flows=2, gin_channels=4)
    F._loaded_fingerprint = ("stale",)
    F.set_option("fp16", 1)
    assert F._options == {"fp16": 1} and F._loaded_fingerprint is None and F._handle is None


def test_zero_padded_channels_compute_the_narrow_decoder():
    """The device carries stages narrower than 32 channels (HiFi-GAN V2) and latents whose width is not a multiple of 32
    zero-padded to 32 (csrc/decoder.cu, Layer::cin_src / cout_src).  The claim behind it, checked on the CPU oracle: with
    zero weights and zero biases on the padding, the padded decoder's output IS the narrow decoder's."""
    hp = oracle.hparams.DecoderHParams(20, "1", (3, 5), ((1, 3, 5), (1, 2, 3)), (4, 2), 32, (8, 4), 0)   # stages 16, 8
    sd = oracle.weights.fold_state_dict(oracle.synth_state_dict(hp, 77, gain=2.0))
    rs = np.random.RandomState(3)
    z = rs.standard_normal((2, hp.initial_channel, 9)).astype(np.float32)
    ref = oracle.generator_forward_np(hp, sd, z, None, dtype=np.float64)

    def pad(a, axes, to=32):
        w = [(0, 0)] * a.ndim
        for ax in axes:
            w[ax] = (0, to - a.shape[ax])
        return np.pad(a, w)

    padded = {}
    for k, v in sd.items():
        if k == "conv_pre.weight":
            padded[k] = pad(v, [1])                 # [c0, initial_channel -> 32, 7]
        elif k == "conv_pre.bias":
            padded[k] = v
        elif k.startswith("ups.0.weight"):
            padded[k] = pad(v, [1])                 # [c_in = c0 = 32, c_out 16 -> 32, k]
        elif k.endswith(".weight"):
            padded[k] = pad(v, [0, 1]) if k != "conv_post.weight" else pad(v, [1])
        else:
            padded[k] = pad(v, [0])                 # biases
    zp = pad(z, [1])
    got = oracle.generator_forward_np(hp, padded, zp, None, dtype=np.float64)
    assert got.shape == ref.shape
    assert np.array_equal(got, ref)                 # exact: the padding contributes exact zeros to every sum


def test_wav_header_is_byte_identical_to_scipy():
    """Output side (cmd_inference.py:117 writes float32 with scipy.io.wavfile.write): the header WavBatchWriter places in
    front of the samples, for the reference's float format and for 16-bit PCM."""
    import io
    import scipy.io.wavfile as wavf
    from importlib import import_module
    wavout = import_module("personalized_text-to-speech_b200.wavout")
    rs = np.random.RandomState(0)
    for n in (0, 1, 7, 22050, 220672):
        x = (rs.standard_normal(n) * 0.3).astype(np.float32)
        for fmt, data in (("float32", x), ("pcm16", wavout.pcm16_reference(x))):
            f = io.BytesIO()
            wavf.write(f, 22050, data)
            want = f.getvalue()
            got = wavout.wav_header(n, 22050, fmt) + data.tobytes()
            assert got == want, (n, fmt)
    assert wavout.pcm16_reference([2.0, -2.0, 0.5, -0.5, 1.0 / 65534])[:4].tolist() == [32767, -32767, 16384, -16384]

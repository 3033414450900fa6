"""GPU: end-to-end parity of vitsdec.Generator (C ABI -> sm_100a kernels) against
  * the committed golden outputs of the unmodified reference (tests/golden/*.npz), and
  * the fp32 torch restatement of the reference at larger sizes,
plus size-independent properties at BASELINE.json's full size.

Stated tolerance (bf16 operands and bf16-stored activations, fp32 accumulation; north_star):
    SNR >= 35 dB   and   max-abs error <= 3 % of the reference waveform's peak.
The reference itself under CPU bf16 autocast sits at 38.5 dB.
"""
import os

import numpy as np
import pytest
import torch

import oracle
import vitsdec
from oracle.generator_torch import generator_forward_torch, to_torch_state_dict
from tests.golden.cases import CASES

pytestmark = pytest.mark.gpu

SNR_MIN_DB = 35.0
MAXABS_FRAC = 0.03
DEV = "cuda:0"


def snr_db(ref, got):
    ref, got = ref.double(), got.double()
    return 10 * np.log10(float((ref ** 2).sum()) / max(float(((ref - got) ** 2).sum()), 1e-300))


def check(ref, got, snr_min=SNR_MIN_DB):
    assert got.shape == ref.shape
    assert torch.isfinite(got).all()
    s = snr_db(ref, got)
    m = float((ref - got).abs().max()) / float(ref.abs().max())
    assert s >= snr_min and m <= MAXABS_FRAC, "SNR %.1f dB, max-abs %.3f of peak" % (s, m)
    return s, m


def build(hp, seed, gain=2.0, impl=0):
    args, kw = hp.ctor_args()
    G = vitsdec.Generator(*args, **kw)
    sd = oracle.synth_state_dict(hp, seed, gain=gain)
    G.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    G = G.to(DEV).eval()
    G.set_option("impl", impl)
    return G, sd


@pytest.mark.parametrize("impl", [0, 1], ids=["tcgen05", "simt"])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_golden_reference_outputs(golden_dir, case, impl):
    name, hp, seed, B, T, use_g = case
    gold = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    G, _ = build(hp, seed, float(gold["gain"]), impl)
    z = torch.from_numpy(gold["z"]).to(DEV)
    g = torch.from_numpy(gold["g"]).to(DEV) if "g" in gold else None
    with torch.no_grad():
        y = G(z, g)
    check(torch.from_numpy(gold["y"]), y.cpu())


@pytest.mark.parametrize("B,T,use_g", [(1, 173, True), (3, 61, True), (2, 100, False), (1, 1, True), (16, 7, True),
                                       (3, 32, True), (3, 32, False)])
def test_full_config_vs_fp32_restatement(B, T, use_g):
    hp = oracle.FINETUNE_SPEAKER
    G, sd = build(hp, 31)
    rs = np.random.RandomState(B * 100 + T)
    z = torch.from_numpy(rs.standard_normal((B, hp.initial_channel, T)).astype(np.float32))
    g = torch.from_numpy(rs.standard_normal((B, hp.gin_channels, 1)).astype(np.float32)) if use_g else None
    ref = generator_forward_torch(hp, to_torch_state_dict(sd), z, g)
    with torch.no_grad():
        y = G(z.to(DEV), None if g is None else g.to(DEV))
    check(ref, y.cpu())


def test_intermediates_track_the_reference_layer_by_layer():
    hp = oracle.FINETUNE_SPEAKER
    G, sd = build(hp, 32)
    G.set_option("debug_keep", 1)
    B, T = 2, 40
    rs = np.random.RandomState(4)
    z = torch.from_numpy(rs.standard_normal((B, hp.initial_channel, T)).astype(np.float32))
    g = torch.from_numpy(rs.standard_normal((B, hp.gin_channels, 1)).astype(np.float32))
    taps = {}
    generator_forward_torch(hp, to_torch_state_dict(sd), z, g, taps=taps)
    with torch.no_grad():
        G(z.to(DEV), g.to(DEV))
    for name, ref in taps.items():
        got = G.debug_read(name, B, T).cpu()
        assert snr_db(ref, got) > 40.0, name


def test_weight_norm_g_is_honoured():
    """At default init g == ||v|| and a fold that ignores g would pass; synth_state_dict perturbs g, and scaling g
    of one layer must change the output."""
    hp = oracle.TINY
    G, sd = build(hp, 33)
    z = torch.randn(1, hp.initial_channel, 8, device=DEV)
    with torch.no_grad():
        y0 = G(z).clone()
        G.ups[0].weight_g.mul_(1.5)
        y1 = G(z)                     # parameter version bump -> refold on the next call
    assert float((y0 - y1).abs().max()) > 1e-4


def test_non_contiguous_slice_half_input_and_reload():
    hp = oracle.TINY
    G, sd = build(hp, 34)
    zfull = torch.randn(2, hp.initial_channel, 40, device=DEV)
    g = torch.randn(2, hp.gin_channels, 1, device=DEV)
    with torch.no_grad():
        a = G(zfull[:, :, :25], g)                      # the slice infer passes (models.py:522)
        b = G(zfull[:, :, :25].contiguous(), g)
        assert torch.equal(a, b)
        h = G(zfull[:, :, :25].half(), g.half())
        assert h.dtype == torch.float16
        G.load_state_dict({k: torch.from_numpy(v) for k, v in oracle.synth_state_dict(hp, 35, gain=2.0).items()})
        c = G(zfull[:, :, :25], g)
    assert float((a - c).abs().max()) > 1e-4            # new weights took effect


def test_full_size_batch_several_utterances_vs_fp32_restatement():
    """BASELINE config 3 size (16 x 862 frames), another seed: utterances 3, 9 and 15 of the batch -- interior and last
    rows of the batch, every tile position of a 10 s utterance -- against the fp32 restatement on the CPU."""
    hp = oracle.FINETUNE_SPEAKER
    G, sd = build(hp, 52)
    B, T = 16, 862
    rs = np.random.RandomState(77)
    z = torch.from_numpy(rs.standard_normal((B, hp.initial_channel, T)).astype(np.float32))
    g = torch.from_numpy(rs.standard_normal((B, hp.gin_channels, 1)).astype(np.float32))
    with torch.no_grad():
        y = G(z.to(DEV), g.to(DEV)).cpu()
    tsd = to_torch_state_dict(sd)
    for u in (3, 9, 15):
        ref = generator_forward_torch(hp, tsd, z[u:u + 1], g[u:u + 1])
        check(ref, y[u:u + 1])


def test_batch_and_determinism_properties_at_full_size():
    """BASELINE config 3 size (16 x 10 s): utterances are independent and the result is reproducible bit for bit;
    utterance 0 is checked against the fp32 restatement on the CPU."""
    hp = oracle.FINETUNE_SPEAKER
    G, sd = build(hp, 36)
    B, T = 16, 862
    rs = np.random.RandomState(8)
    z = torch.from_numpy(rs.standard_normal((B, hp.initial_channel, T)).astype(np.float32)).to(DEV)
    g = torch.from_numpy(rs.standard_normal((B, hp.gin_channels, 1)).astype(np.float32)).to(DEV)
    with torch.no_grad():
        y = G(z, g)
        y2 = G(z, g)
        y0 = G(z[:1], g[:1])
        y7 = G(z[7:9], g[7:9])
    assert y.shape == (B, 1, T * 256) and torch.isfinite(y).all() and float(y.abs().max()) <= 1.0
    assert torch.equal(y, y2)
    assert torch.equal(y[:1], y0) and torch.equal(y[7:9], y7)
    ref0 = generator_forward_torch(hp, to_torch_state_dict(sd), z[:1].cpu(), g[:1].cpu())
    check(ref0, y0.cpu())


def test_chunked_long_form_matches_unchunked():
    """BASELINE config 5 (uma_trilingual hyper-parameters == finetune_speaker's): chunks + 12-frame halos."""
    hp = oracle.UMA_TRILINGUAL
    G, sd = build(hp, 37)
    T = 1300
    z = torch.randn(1, hp.initial_channel, T, device=DEV)
    g = torch.randn(1, hp.gin_channels, 1, device=DEV)
    with torch.no_grad():
        full = G(z, g)
        chunked = vitsdec.decode_chunked(G, z, g, chunk_frames=256, halo=12, hop=hp.hop)
    assert chunked.shape == full.shape
    # The receptive field is covered by the halo.  What is left is bf16 re-rounding: a chunk starts at another phase of
    # the time-folded tiles, so its fp32 sums are formed in another order than the unchunked decode's.
    assert snr_db(full.cpu(), chunked.cpu()) > 42.0
    ref = generator_forward_torch(hp, to_torch_state_dict(sd), z.cpu(), g.cpu())
    check(ref, chunked.cpu())
    check(ref, full.cpu())
    G.set_option("fold", 0)   # plain tiles sum in a position-independent order: chunks reproduce the full decode
    with torch.no_grad():
        full0 = G(z, g)
        chunked0 = vitsdec.decode_chunked(G, z, g, chunk_frames=256, halo=12, hop=hp.hop)
    assert snr_db(full0.cpu(), chunked0.cpu()) > 60.0


def test_decode_host_entry_matches_device_entry():
    import ctypes
    hp = oracle.TINY
    G, sd = build(hp, 38)
    B, T = 2, 30
    z = torch.randn(B, hp.initial_channel, T)
    g = torch.randn(B, hp.gin_channels, 1)
    with torch.no_grad():
        y = G(z.to(DEV), g.to(DEV)).cpu()
    out = torch.empty(B, 1, T * hp.hop)
    lib = vitsdec._capi.lib()
    vitsdec._capi.check(lib.vitsdec_decode_host(G._handle, z.data_ptr(), g.data_ptr(), out.data_ptr(), B, T))
    assert torch.equal(out, y)


def test_sharded_single_rank_is_identity():
    hp = oracle.TINY
    G, sd = build(hp, 39)
    z = torch.randn(3, hp.initial_channel, 12, device=DEV)
    g = torch.randn(3, hp.gin_channels, 1, device=DEV)
    with torch.no_grad():
        assert torch.equal(vitsdec.decode_sharded(G, z, g), G(z, g))


def test_fused_pairs_equal_unfused_schedule():
    hp = oracle.FINETUNE_SPEAKER
    G, sd = build(hp, 40)
    B, T = 2, 150
    z = torch.randn(B, hp.initial_channel, T, device=DEV)
    g = torch.randn(B, hp.gin_channels, 1, device=DEV)
    G.set_option("fold", 0)   # the time-folded forms sum in a different order; compare like with like:
    G.set_option("mrfp", 0)   # conv_pair.cu (plain tiles, taps in order) against one launch per conv
    with torch.no_grad():
        a = G(z, g)
        n_fused = G.last_launch_count()
        G.set_option("fuse_pairs", 0)
        b = G(z, g)
        n_plain = G.last_launch_count()
        G.set_option("fuse_pairs", 1)
        G.set_option("mrfp", 1)   # default: the C = 32 pairs on the 2-sample folded view (conv_mrfp.cu), another tap order
        c = G(z, g)
    assert n_fused < n_plain
    assert torch.equal(a, b)
    assert snr_db(a.cpu(), c.cpu()) > 45.0


def test_time_folded_schedule_tracks_the_unfolded_one():
    """Narrow dilation-1 layers run time-folded by default; option fold=0 is the plain schedule."""
    hp = oracle.FINETUNE_SPEAKER
    G, sd = build(hp, 41)
    B, T = 2, 173
    z = torch.randn(B, hp.initial_channel, T, device=DEV)
    g = torch.randn(B, hp.gin_channels, 1, device=DEV)
    with torch.no_grad():
        a = G(z, g).clone()
        G.set_option("fold", 0)
        b = G(z, g).clone()
        G.set_option("fuse_pairs", 0)
        G.set_option("fold", 1)
        c = G(z, g).clone()
        G.set_option("fuse_pairs", 1)
        G.set_option("pairf", 2)   # the folded fused-pair kernel wherever it exists, not only where it is faster
        e = G(z, g).clone()
    ref = generator_forward_torch(hp, to_torch_state_dict(sd), z.cpu(), g.cpu())
    for y in (a, b, c, e):
        check(ref, y.cpu())
    assert snr_db(b.cpu(), a.cpu()) > 45.0 and snr_db(b.cpu(), c.cpu()) > 45.0 and snr_db(b.cpu(), e.cpu()) > 45.0


def test_large_batch_matches_small_batches_bitwise():
    """BASELINE config 4's per-GPU shard on two GPUs (128 x 10 s = 18 GB of workspace): index arithmetic at 2^31-scale
    element counts.  Utterances are independent, so slices of the big batch equal the same utterances decoded alone."""
    hp = oracle.FINETUNE_SPEAKER
    G, sd = build(hp, 42)
    B, T = 128, 862
    gen = torch.Generator(device=DEV).manual_seed(9)
    z = torch.randn(B, hp.initial_channel, T, device=DEV, generator=gen)
    g = torch.randn(B, hp.gin_channels, 1, device=DEV, generator=gen)
    with torch.no_grad():
        y = G(z, g)
        assert y.shape == (B, 1, T * 256) and torch.isfinite(y).all()
        for lo in (0, 63, 127):
            assert torch.equal(y[lo:lo + 1], G(z[lo:lo + 1], g[lo:lo + 1]))
    del y
    torch.cuda.empty_cache()


def test_sixty_second_utterance_chunked_at_config5_size():
    """BASELINE config 5: one 60 s utterance (T = 5168), 512-frame chunks with 12-frame halos vs the unchunked decode."""
    hp = oracle.UMA_TRILINGUAL
    G, sd = build(hp, 43)
    T = 5168
    z = torch.randn(1, hp.initial_channel, T, device=DEV)
    g = torch.randn(1, hp.gin_channels, 1, device=DEV)
    with torch.no_grad():
        full = G(z, g)
        chunked = vitsdec.decode_chunked(G, z, g, chunk_frames=512, halo=12, hop=hp.hop)
    assert chunked.shape == full.shape == (1, 1, T * 256)
    assert snr_db(full.cpu(), chunked.cpu()) > 42.0
    # utterance-level check against the fp32 restatement on a window (the full 60 s takes minutes on the CPU)
    w0, w1 = 2000, 2200
    ref = generator_forward_torch(hp, to_torch_state_dict(sd), z[:, :, w0 - 12:w1 + 12].cpu(), g.cpu())
    check(ref[:, :, 12 * 256:-12 * 256], chunked[:, :, w0 * 256:w1 * 256].cpu())


def test_concurrent_forward_from_worker_threads():
    """Gradio calls tts_fn from worker threads (VC_inference.py:38-53): forward on one module from several threads, each
    on its own stream, must give the single-threaded results (per-call workspace, plan building behind a mutex)."""
    import threading
    hp = oracle.FINETUNE_SPEAKER
    G, sd = build(hp, 44)
    shapes = [(1, 40), (2, 33), (1, 40), (3, 21)]
    zs = [torch.randn(b, hp.initial_channel, t, device=DEV) for b, t in shapes]
    gs = [torch.randn(b, hp.gin_channels, 1, device=DEV) for b, _ in shapes]
    with torch.no_grad():
        want = [G(z, g).clone() for z, g in zip(zs, gs)]
    torch.cuda.synchronize()
    got = [None] * len(shapes)
    errs = []

    def work(i):
        try:
            s = torch.cuda.Stream(device=DEV)
            with torch.no_grad(), torch.cuda.stream(s):
                for _ in range(5):
                    y = G(zs[i], gs[i])
                got[i] = y.clone()
            s.synchronize()
        except Exception as e:  # surfaced below
            errs.append(e)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(shapes))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    torch.cuda.synchronize()
    assert not errs, errs
    for w, y in zip(want, got):
        assert torch.equal(w, y)


def test_concurrent_branches_equal_the_serial_schedule():
    """Option par: the MRF branches of a stage on forked streams under the CUDA graph (default on)."""
    hp = oracle.FINETUNE_SPEAKER
    G, sd = build(hp, 45)
    for B, T in ((1, 173), (2, 600)):
        z = torch.randn(B, hp.initial_channel, T, device=DEV)
        g = torch.randn(B, hp.gin_channels, 1, device=DEV)
        outs = []
        for par in (0, 1):
            G.set_option("par", par)
            with torch.no_grad():
                for _ in range(4):             # plain launches twice, then graph capture, then graph replay
                    outs.append(G(z, g).clone())
        for y in outs[1:]:
            assert torch.equal(outs[0], y)


def test_host_pipeline_matches_direct_forward():
    """vitsdec.HostPipeline: pinned host buffers in, pinned waveforms out, several batches in flight."""
    hp = oracle.FINETUNE_SPEAKER
    G, sd = build(hp, 46)
    B, T = 2, 90
    zs = [torch.randn(B, hp.initial_channel, T).pin_memory() for _ in range(5)]
    gs = [torch.randn(B, hp.gin_channels, 1).pin_memory() for _ in range(5)]
    outs = [torch.empty(B, 1, T * hp.hop).pin_memory() for _ in range(5)]
    pipe = vitsdec.HostPipeline(G, depth=2)
    for z, g, o in zip(zs, gs, outs):
        pipe.submit(z, g, o)
    pipe.wait_all()
    with torch.no_grad():
        for z, g, o in zip(zs, gs, outs):
            assert torch.equal(o, G(z.to(DEV), g.to(DEV)).cpu())
    with pytest.raises(RuntimeError, match="pinned"):
        pipe.submit(torch.randn(B, hp.initial_channel, T), gs[0], outs[0])
    # on_device hook (the decoded batch still on the GPU, in the slot's stream) and decode_after (the next decode waits
    # for an event recorded behind a side-stream consumer -- how bench.py orders an NCCL gather between decodes)
    side = torch.cuda.Stream(device=DEV)
    seen, last = [], [None]

    def hook(y, stream):
        side.wait_stream(stream)
        with torch.cuda.stream(side):
            seen.append(y.sum())
            y.record_stream(side)
            ev = torch.cuda.Event()
            ev.record(side)
        last[0] = ev

    for z, g, o in zip(zs, gs, outs):
        o.zero_()
        pipe.submit(z, g, o, on_device=hook, decode_after=last[0])
    pipe.wait_all()
    torch.cuda.synchronize()
    with torch.no_grad():
        for z, g, o, sm in zip(zs, gs, outs, seen):
            ref = G(z.to(DEV), g.to(DEV))
            assert torch.equal(o, ref.cpu()) and torch.equal(sm, ref.sum())


# ---- option "fp16": fp16 instead of bf16 operands / stored activations (north_star: "bf16/fp16 operands, fp32
# accumulation").  Stated tolerance for this mode: SNR >= 50 dB and max-abs error <= 0.5 % of the reference peak
# (three more mantissa bits per stored activation = 18 dB over the bf16 mode's measured 39 dB).
FP16_SNR_MIN_DB = 50.0
FP16_MAXABS_FRAC = 0.005


def check_fp16(ref, got):
    assert got.shape == ref.shape and torch.isfinite(got).all()
    s = snr_db(ref, got)
    m = float((ref - got).abs().max()) / float(ref.abs().max())
    assert s >= FP16_SNR_MIN_DB and m <= FP16_MAXABS_FRAC, "SNR %.1f dB, max-abs %.4f of peak" % (s, m)
    return s, m


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_fp16_mode_golden_reference_outputs(golden_dir, case):
    name, hp, seed, B, T, use_g = case
    gold = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    G, _ = build(hp, seed, float(gold["gain"]))
    G.set_option("fp16", 1)
    z = torch.from_numpy(gold["z"]).to(DEV)
    g = torch.from_numpy(gold["g"]).to(DEV) if "g" in gold else None
    with torch.no_grad():
        y = G(z, g)
    check_fp16(torch.from_numpy(gold["y"]), y.cpu())


@pytest.mark.parametrize("B,T", [(1, 173), (3, 61), (16, 7)])
def test_fp16_mode_vs_fp32_restatement_and_back(B, T):
    hp = oracle.FINETUNE_SPEAKER
    G, sd = build(hp, 31)
    rs = np.random.RandomState(B * 100 + T)
    z = torch.from_numpy(rs.standard_normal((B, hp.initial_channel, T)).astype(np.float32))
    g = torch.from_numpy(rs.standard_normal((B, hp.gin_channels, 1)).astype(np.float32))
    ref = generator_forward_torch(hp, to_torch_state_dict(sd), z, g)
    with torch.no_grad():
        y_bf = G(z.to(DEV), g.to(DEV)).clone()
        G.set_option("fp16", 1)                 # with a live native handle: weights are re-folded as fp16
        y_h = G(z.to(DEV), g.to(DEV)).clone()
        G.set_option("fp16", 0)
        y_bf2 = G(z.to(DEV), g.to(DEV))
    s_bf, _ = check(ref, y_bf.cpu())
    s_h, _ = check_fp16(ref, y_h.cpu())
    assert s_h > s_bf + 10.0
    assert torch.equal(y_bf, y_bf2)


def test_fp16_mode_schedules_agree():
    """fused pairs vs single launches (bit-identical on plain tiles), folded vs plain tiles, tcgen05 vs CUDA cores."""
    hp = oracle.FINETUNE_SPEAKER
    G, sd = build(hp, 36)
    G.set_option("fp16", 1)
    z = torch.randn(2, hp.initial_channel, 64, device=DEV)
    g = torch.randn(2, hp.gin_channels, 1, device=DEV)
    with torch.no_grad():
        y = G(z, g).clone()
        G.set_option("fold", 0)
        G.set_option("mrfp", 0)   # plain-tile pairs (conv_pair.cu): the like-for-like partner of single launches
        a = G(z, g).clone()
        G.set_option("fuse_pairs", 0)
        b = G(z, g).clone()
        G.set_option("impl", 1)
        c = G(z, g).clone()
    assert torch.equal(a, b)
    assert snr_db(a.cpu(), y.cpu()) > 60.0
    assert snr_db(c.cpu(), a.cpu()) > 60.0


def test_fp16_mode_full_size_determinism():
    hp = oracle.FINETUNE_SPEAKER
    G, sd = build(hp, 37)
    G.set_option("fp16", 1)
    torch.manual_seed(5)
    z = torch.randn(16, hp.initial_channel, 862, device=DEV)
    g = torch.randn(16, hp.gin_channels, 1, device=DEV)
    with torch.no_grad():
        y0 = G(z, g).clone()
        y1 = G(z, g).clone()
        y_one = G(z[5:6], g[5:6])
    assert torch.isfinite(y0).all() and torch.equal(y0, y1)
    assert torch.equal(y0[5:6], y_one)       # batch independence, bit-exact
    ref = generator_forward_torch(hp, to_torch_state_dict(sd), z[5:6].cpu(), g[5:6].cpu())
    check_fp16(ref, y_one.cpu())


# ---- other HiFi-GAN configurations through the same constructor
HIFIGAN_V3_LIKE = oracle.hparams.DecoderHParams(80, "2", (3, 5, 7), ((1, 2), (2, 6), (3, 12)), (8, 8, 4), 256, (16, 16, 8), 0)
WIDE_K13 = oracle.hparams.DecoderHParams(96, "1", (3, 13), ((1, 3, 5), (1, 2, 4)), (5, 4, 2), 256, (11, 8, 4), 64)


@pytest.mark.parametrize("hp,B,T,fp16", [(HIFIGAN_V3_LIKE, 2, 50, 0), (HIFIGAN_V3_LIKE, 1, 211, 1), (WIDE_K13, 3, 33, 0),
                                         (WIDE_K13, 1, 90, 1)],
                         ids=["v3like_bf16", "v3like_fp16", "k13_stride5_bf16", "k13_stride5_fp16"])
def test_other_hifigan_configs_vs_fp32_restatement(hp, B, T, fp16):
    """ResBlock2 with the published HiFi-GAN V3 shape (3 stages, rates 8/8/4, dilations up to 12, no speaker input) and a
    ResBlock1 variant with an odd stride (5, kernel 11), k = 13 and 2 kernels per stage: nothing in the kernels is
    specific to configs/finetune_speaker.json."""
    G, sd = build(hp, 91)
    G.set_option("fp16", fp16)
    rs = np.random.RandomState(B * 10 + T)
    z = torch.from_numpy(rs.standard_normal((B, hp.initial_channel, T)).astype(np.float32))
    g = torch.from_numpy(rs.standard_normal((B, hp.gin_channels, 1)).astype(np.float32)) if hp.gin_channels else None
    ref = generator_forward_torch(hp, to_torch_state_dict(sd), z, g)
    with torch.no_grad():
        y = G(z.to(DEV), None if g is None else g.to(DEV))
    assert y.shape == (B, 1, T * hp.hop)
    (check_fp16 if fp16 else check)(ref, y.cpu())


HIFIGAN_V2_LIKE = oracle.hparams.DecoderHParams(80, "1", (3, 7, 11), ((1, 3, 5),) * 3, (8, 8, 2, 2), 128, (16, 16, 4, 4), 0)


@pytest.mark.parametrize("fp16", [0, 1], ids=["bf16", "fp16"])
def test_hifigan_v2_narrow_stages(fp16):
    """HiFi-GAN V2 narrows to 16 and 8 channels: those stages are carried zero-padded to the 32-channel granularity of
    the tiles (zero weights and biases keep the padding at zero), so the result is the narrow decoder's."""
    hp = HIFIGAN_V2_LIKE
    G, sd = build(hp, 92)
    G.set_option("fp16", fp16)
    rs = np.random.RandomState(17)
    z = torch.from_numpy(rs.standard_normal((2, hp.initial_channel, 37)).astype(np.float32))
    ref = generator_forward_torch(hp, to_torch_state_dict(sd), z, None)
    with torch.no_grad():
        y = G(z.to(DEV))
        G.set_option("fold", 0)
        G.set_option("fuse_pairs", 0)
        y_plain = G(z.to(DEV))
    (check_fp16 if fp16 else check)(ref, y.cpu())
    (check_fp16 if fp16 else check)(ref, y_plain.cpu())


def test_unsupported_widths_fail_loudly():
    """upsample_initial_channel itself must be a multiple of 32: a clear error, no fallback."""
    hp = oracle.hparams.DecoderHParams(80, "1", (3,), ((1, 3, 5),), (2,), 48, (4,), 0)
    args, kw = hp.ctor_args()
    G = vitsdec.Generator(*args, **kw).to(DEV).eval()
    with pytest.raises(Exception, match="multiple of 32"):
        with torch.no_grad():
            G(torch.randn(1, 80, 8, device=DEV))


def test_programmatic_dependent_launch_modes_agree():
    """Option "pdl": short launches start their prologue under the previous launch of the stream (griddepcontrol); the
    result must not depend on it -- plain launches (first uses of a plan) and the captured graph (third use on)."""
    hp = oracle.FINETUNE_SPEAKER
    G, sd = build(hp, 38)
    outs = {}
    for B, T in ((1, 40), (2, 173), (3, 300)):
        z = torch.randn(B, hp.initial_channel, T, device=DEV)
        g = torch.randn(B, hp.gin_channels, 1, device=DEV)
        for pdl in (0, 1, 2):
            G.set_option("pdl", pdl)
            with torch.no_grad():
                ys = [G(z, g).clone() for _ in range(5)]
            assert all(torch.equal(ys[0], y) for y in ys[1:])
            outs[pdl] = ys[0]
        assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    ref = generator_forward_torch(hp, to_torch_state_dict(sd), z.cpu(), g.cpu())
    check(ref, outs[1].cpu())


def test_wav_output_bytes_equal_scipy_write(tmp_path):
    """SURVEY.md 8f-2, the output side of cmd_inference.py:114-117: WavBatchWriter (device-side PCM16 conversion, async
    D2H straight behind the header in pinned memory) writes files whose bytes equal scipy.io.wavfile.write of the same
    samples -- float32 (the format the reference writes) and pcm16 -- and whose content matches the fp32 reference
    decode within the decoder's tolerance."""
    import io
    import scipy.io.wavfile as wavf
    from importlib import import_module
    wavout = import_module("personalized_text-to-speech_b200.wavout")
    hp = oracle.FINETUNE_SPEAKER
    G, sd = build(hp, 77)
    rs = np.random.RandomState(12)
    B, T = 3, 40
    z = torch.from_numpy(rs.standard_normal((B, hp.initial_channel, T)).astype(np.float32))
    g = torch.from_numpy(rs.standard_normal((B, hp.gin_channels, 1)).astype(np.float32))
    ref = generator_forward_torch(hp, to_torch_state_dict(sd), z, g)
    with torch.no_grad():
        y = G(z.to(DEV), g.to(DEV)) * 30.0        # random-init waveforms peak at ~0.03: bring them to PCM scale
    ref = ref * 30.0
    lengths = [T * 256, 31 * 256 + 5, 1]
    for fmt in ("float32", "pcm16"):
        w = vitsdec.WavBatchWriter(22050, fmt)
        images, ev = w.enqueue(y, lengths)
        paths = [str(tmp_path / ("u%d_%s.wav" % (i, fmt))) for i in range(B)]
        w.save(images, paths, ev)
        for i, path in enumerate(paths):
            mine = y[i, 0, :lengths[i]].cpu().numpy()
            data = mine if fmt == "float32" else wavout.pcm16_reference(mine)
            f = io.BytesIO()
            wavf.write(f, 22050, data)
            assert open(path, "rb").read() == f.getvalue(), (fmt, i)
            sr, back = wavf.read(path)
            assert sr == 22050 and back.shape == (lengths[i],)
            want = ref[i, 0, :lengths[i]].numpy()
            if fmt == "pcm16":   # vs the REFERENCE output: within the decoder's 3 %-of-peak bound, in PCM steps
                want16 = wavout.pcm16_reference(want).astype(np.int64)
                peak = max(1, int(np.abs(wavout.pcm16_reference(ref.numpy())).max()))
                assert np.abs(back.astype(np.int64) - want16).max() <= 0.03 * peak + 1
            elif lengths[i] > 256:
                check(torch.from_numpy(want), torch.from_numpy(back))


def test_default_path_keeps_plans_and_reports_graph_failures():
    """the default forward must not rebuild launch plans when the allocator moves the workspace -- the module
    keeps its workspace per stream -- and a plan that fell back from its CUDA graph must be visible."""
    hp = oracle.FINETUNE_SPEAKER
    G, sd = build(hp, 78)
    rs = np.random.RandomState(3)
    z = torch.from_numpy(rs.standard_normal((1, hp.initial_channel, 50)).astype(np.float32)).to(DEV)
    g = torch.from_numpy(rs.standard_normal((1, hp.gin_channels, 1)).astype(np.float32)).to(DEV)
    with torch.no_grad():
        y0 = G(z, g)
        ptr = next(iter(G._ws.values())).data_ptr()
        junk = [torch.empty(1 << 20, device=DEV) for _ in range(8)]   # allocator churn between calls
        for _ in range(6):
            y = G(z, g)
            assert next(iter(G._ws.values())).data_ptr() == ptr
            assert torch.equal(y, y0)
        del junk
        # mixed shapes share the workspace; an in-place parameter update is still noticed without assume_frozen
        y_small = G(z[:, :, :20].contiguous(), g)
        assert y_small.shape == (1, 1, 20 * 256)
        G.ups[0].weight_g.mul_(1.25)
        y2 = G(z, g)
        assert not torch.equal(y2, y0)
        G.ups[0].weight_g.div_(1.25)
    assert G.get_option("graph_failed") == 0
    assert G.get_option("testing_build") == 0    # the product library has no second backend
    with pytest.raises(Exception):
        vitsdec._capi.check(vitsdec._capi.lib().vitsdec_set_option(G._handle, b"impl", 1), "set_option")


@pytest.mark.parametrize("B,T", [(1, 1), (1, 7), (3, 32), (2, 100), (16, 20)])
def test_fused_last_pairs_launch_matches_separate_launches(B, T):
    """conv_mrfp.cu (option "mrfp", default on): every ResBlock pair of the C = 32 stage on the 2-sample folded view, and
    the last pair of every MRF branch, the branch sum and the average in ONE launch.  Against the schedule it replaces
    (conv_pair.cu pairs, three c1 launches + the fused-MRF launch): within bf16 re-rounding of each other (the folded
    kernels accumulate the taps in another order), both inside the stated tolerance of the fp32 restatement.  Three
    launches fewer per decode."""
    hp = oracle.FINETUNE_SPEAKER
    G, sd = build(hp, 41)
    rs = np.random.RandomState(B * 31 + T)
    z = torch.from_numpy(rs.standard_normal((B, hp.initial_channel, T)).astype(np.float32))
    g = torch.from_numpy(rs.standard_normal((B, hp.gin_channels, 1)).astype(np.float32))
    ref = generator_forward_torch(hp, to_torch_state_dict(sd), z, g)
    out = {}
    with torch.no_grad():
        for fold in (0, 1):
            for mrfp in (0, 1, 3):   # 3: the C = 64 pairs through the same kernel as well (plain rows)
                G.set_option("fold", fold)
                G.set_option("mrfp", mrfp)
                out[fold, mrfp] = G(z.to(DEV), g.to(DEV)).cpu()
                out["n", fold, mrfp] = G.last_launch_count()
    assert snr_db(out[0, 0], out[0, 1]) > 45.0
    assert snr_db(out[1, 0], out[1, 1]) > 45.0
    assert snr_db(out[1, 0], out[1, 3]) > 45.0 and out["n", 1, 3] == out["n", 1, 1]
    assert out["n", 1, 1] == out["n", 1, 0] - 3 and out["n", 0, 1] == out["n", 0, 0] - 3
    if T >= 7:
        for k in ((0, 1), (1, 1)):
            check(ref, out[k])


def test_peer_memory_gather_single_rank(tmp_path):
    """sharding.PeerGather on a one-rank group (the driver's GPU test tier has one GPU): the symmetric buffer, the
    copy-engine push behind a producer stream and the closing barrier.  The 2 / 4 / 8-rank path is what bench.py's e2e leg
    runs under torchrun."""
    import torch.distributed as dist
    if dist.is_initialized():
        pytest.skip("a process group already exists in this process")
    try:
        dist.init_process_group("nccl", init_method="file://%s" % (tmp_path / "pg"), rank=0, world_size=1,
                                device_id=torch.device(DEV))
    except Exception as e:   # pragma: no cover - environment without NCCL
        pytest.skip("no NCCL process group: %r" % (e,))
    try:
        try:
            pg = vitsdec.PeerGather((4, 1, 2560), torch.float32, torch.device(DEV))
        except Exception as e:
            pytest.skip("symmetric memory unavailable: %r" % (e,))
        s = torch.cuda.Stream(device=DEV)
        with torch.cuda.stream(s):
            y0 = torch.randn(4, 1, 2560, device=DEV)
            y1 = torch.randn(4, 1, 2560, device=DEV)
            pg.push(y0, 0, after=s)
            pg.push(y1, 1, after=s)
        pg.finish()
        torch.cuda.synchronize()
        assert torch.equal(pg.full(0), y0) and torch.equal(pg.full(1), y1)
    finally:
        dist.destroy_process_group()

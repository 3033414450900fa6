"""GPU: per-kernel parity of the fused convolution primitives against torch functional ops on bf16-rounded
operands (SURVEY.md section 4 'per-kernel').  Calls go through the C ABI (vitsdec_op_*)."""
import importlib

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

ops = importlib.import_module("personalized_text-to-speech_b200.ops")

# bf16 output rounding is 2^-9 relative; operands are identical bf16 values on both sides
REL_TOL = 1e-2


def ref_conv(x, w, bias, dilation, res, res_gain, out_slope):
    k = w.shape[2]
    y = F.conv1d(x.float().transpose(1, 2), w.bfloat16().float(), bias, dilation=dilation,
                 padding=(k - 1) // 2 * dilation)
    if res is not None:
        r = res.float().transpose(1, 2)
        y = y + torch.where(r >= 0, r, r * res_gain)
    return torch.where(y >= 0, y, y * out_slope).transpose(1, 2)


def rel_err(a, b):
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


CASES = [
    # B, L, c_in, c_out, k, dil, residual
    (1, 128, 64, 32, 1, 1, False),      # plain GEMM tile, every BN instance
    (1, 128, 64, 64, 1, 1, False),
    (1, 256, 64, 128, 1, 1, False),
    (1, 128, 64, 256, 1, 1, False),
    (1, 128, 32, 32, 1, 1, False),      # 64-byte rows (SWIZZLE_64B)
    (1, 128, 64, 64, 3, 8, False),      # tap shift = whole swizzle atoms
    (1, 128, 64, 64, 3, 1, False),      # tap shift inside a swizzle atom
    (2, 300, 128, 128, 7, 5, True),     # stage-1 shape, dilation 5, residual
    (2, 1000, 256, 256, 11, 5, True),   # stage-0 shape, widest halo (50 rows)
    (3, 777, 32, 32, 11, 3, True),      # stage-3 shape, ragged length
    (2, 517, 64, 64, 7, 3, True),       # stage-2 shape
    (2, 50, 192, 512, 7, 1, False),     # conv_pre shape (3 K-chunks, 2 N-tiles)
    (1, 5, 64, 64, 7, 1, True),         # shorter than the kernel's halo
    (1, 1, 64, 64, 11, 5, True),        # single row
    (1, 4000, 64, 64, 11, 5, True),     # many M-tiles on one utterance
    (5, 130, 96, 96, 3, 1, False),      # channels not a power of two (BN=32, KC=32)
]


@pytest.mark.parametrize("impl", [0, 1], ids=["tcgen05", "simt"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "B%d_L%d_ci%d_co%d_k%d_d%d_r%d" % c)
def test_conv1d_matches_torch(case, impl):
    B, L, ci, co, k, d, use_res = case
    torch.manual_seed(B * 1000 + L)
    dev = torch.device("cuda:0")
    x = torch.randn(B, L, ci, device=dev).bfloat16()
    w = torch.randn(co, ci, k, device=dev) / (ci * k) ** 0.5
    b = torch.randn(co, device=dev) * 0.1
    res = torch.randn(B, L, co, device=dev).bfloat16() if use_res else None
    y = ops.conv1d_cl(x, w, b, dilation=d, res=res, res_gain=10.0, out_slope=0.1, impl=impl)
    torch.cuda.synchronize()
    assert rel_err(y, ref_conv(x, w, b, d, res, 10.0, 0.1)) < REL_TOL


def test_tcgen05_equals_simt_up_to_accumulation_order():
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    x = torch.randn(2, 700, 128, device=dev).bfloat16()
    w = torch.randn(128, 128, 7, device=dev) / 30
    b = torch.randn(128, device=dev) * 0.1
    a = ops.conv1d_cl(x, w, b, dilation=3, out_slope=0.1, impl=0).float()
    c = ops.conv1d_cl(x, w, b, dilation=3, out_slope=0.1, impl=1).float()
    # identical bf16 operands, fp32 accumulation in a different order: at most one bf16 ulp apart
    assert float(((a - c).abs() / (c.abs() + 1e-3)).max()) < 2 ** -7


@pytest.mark.parametrize("impl", [0, 1], ids=["tcgen05", "simt"])
@pytest.mark.parametrize("case", [(2, 50, 64, 32, 4, 2), (1, 37, 128, 64, 4, 2), (2, 33, 512, 256, 16, 8),
                                  (1, 130, 256, 128, 16, 8), (1, 9, 64, 32, 8, 4), (1, 1, 64, 32, 4, 2),
                                  (2, 21, 64, 32, 6, 2)],
                         ids=lambda c: "B%d_L%d_ci%d_co%d_k%d_s%d" % c)
def test_conv_transpose1d_matches_torch(case, impl):
    B, L, ci, co, k, s = case
    torch.manual_seed(L)
    dev = torch.device("cuda:0")
    x = torch.randn(B, L, ci, device=dev).bfloat16()
    w = torch.randn(ci, co, k, device=dev) / (ci * k / s) ** 0.5
    b = torch.randn(co, device=dev) * 0.1
    y = ops.conv_transpose1d_cl(x, w, b, stride=s, out_slope=0.1, impl=impl)
    torch.cuda.synchronize()
    r = F.conv_transpose1d(x.float().transpose(1, 2), w.bfloat16().float(), b, stride=s, padding=(k - s) // 2)
    r = torch.where(r >= 0, r, r * 0.1).transpose(1, 2)
    assert y.shape == r.shape
    assert rel_err(y, r) < REL_TOL


def test_streamed_and_resident_weight_paths_agree_bitwise():
    """desc_mode bit 1 forces the streamed-weights pipeline where the planner would keep weights resident."""
    torch.manual_seed(3)
    dev = torch.device("cuda:0")
    for (ci, co, k, d) in ((32, 32, 11, 5), (64, 64, 7, 3), (128, 128, 3, 1)):
        x = torch.randn(2, 1500, ci, device=dev).bfloat16()
        w = torch.randn(co, ci, k, device=dev) / (ci * k) ** 0.5
        b = torch.randn(co, device=dev) * 0.1
        a = ops.conv1d_cl(x, w, b, dilation=d, out_slope=0.1, impl=0, desc_mode=0)
        c = ops.conv1d_cl(x, w, b, dilation=d, out_slope=0.1, impl=0, desc_mode=2)
        assert torch.equal(a, c)


def test_channels_as_m_and_time_as_m_forms_agree_bitwise():
    """desc_mode bit 3 keeps wide layers on the time-as-M tile; the default for them is channels-as-M (SWAP)."""
    torch.manual_seed(4)
    dev = torch.device("cuda:0")
    for (ci, co, k, d, use_res) in ((128, 128, 7, 3, True), (256, 256, 11, 5, True), (192, 512, 7, 1, False),
                                    (256, 256, 3, 1, False)):
        x = torch.randn(2, 1111, ci, device=dev).bfloat16()
        w = torch.randn(co, ci, k, device=dev) / (ci * k) ** 0.5
        b = torch.randn(co, device=dev) * 0.1
        res = torch.randn(2, 1111, co, device=dev).bfloat16() if use_res else None
        a = ops.conv1d_cl(x, w, b, dilation=d, res=res, out_slope=0.1, impl=0, desc_mode=0)
        c = ops.conv1d_cl(x, w, b, dilation=d, res=res, out_slope=0.1, impl=0, desc_mode=8)
        assert torch.equal(a, c)
        assert rel_err(a, ref_conv(x, w, b, d, res, 10.0, 0.1)) < REL_TOL


def test_paired_tiles_agree_bitwise_with_single_cta_tiles():
    """256-output-channel layers run as tcgen05.mma.cta_group::2 tiles across a 2-CTA cluster (conv_tc2.cu) once the
    problem is large enough for 256-row tiles; desc_mode bit 12 keeps them on conv_tc.cu's single-CTA tiles, bit 13 selects
    the relay variant of the operand barriers.  Same products, same accumulation order: bit-identical.  Ragged lengths put a
    partial tile (and a peer CTA whose whole half lies past the utterance end) at the end of every utterance."""
    torch.manual_seed(14)
    dev = torch.device("cuda:0")
    for (B, L, k, d, use_res) in ((5, 2048, 11, 5, True), (7, 1600 + 37, 7, 3, True), (6, 1793, 3, 1, False),
                                  (16, 6896, 11, 1, True)):
        x = torch.randn(B, L, 256, device=dev).bfloat16()
        w = torch.randn(256, 256, k, device=dev) / (256 * k) ** 0.5
        b = torch.randn(256, device=dev) * 0.1
        res = torch.randn(B, L, 256, device=dev).bfloat16() if use_res else None
        a = ops.conv1d_cl(x, w, b, dilation=d, res=res, out_slope=0.1, impl=0, desc_mode=0)
        c = ops.conv1d_cl(x, w, b, dilation=d, res=res, out_slope=0.1, impl=0, desc_mode=4096)
        r = ops.conv1d_cl(x, w, b, dilation=d, res=res, out_slope=0.1, impl=0, desc_mode=8192)
        assert torch.equal(a, c) and torch.equal(a, r)
        if B * L < 20000:
            assert rel_err(a, ref_conv(x, w, b, d, res, 10.0, 0.1)) < REL_TOL
    xh = torch.randn(5, 2048, 256, device=dev).half()      # fp16 storage instance
    w = torch.randn(256, 256, 7, device=dev) / (256 * 7) ** 0.5
    assert torch.equal(ops.conv1d_cl(xh, w, b, out_slope=0.1, desc_mode=0), ops.conv1d_cl(xh, w, b, out_slope=0.1, desc_mode=4096))


@pytest.mark.parametrize("case", [(2, 1024, 32, 3, True), (3, 1500, 32, 7, False), (2, 3108, 32, 11, True),
                                  (2, 1000, 64, 3, True), (1, 518, 64, 7, False), (2, 2222, 64, 11, True),
                                  (1, 4, 32, 11, True), (1, 2, 64, 3, False), (16, 260, 32, 7, True)],
                         ids=lambda c: "B%d_L%d_C%d_k%d_r%d" % c)
def test_time_folded_conv_matches_torch_and_unfolded(case):
    """desc_mode bit 4: the same dilation-1 conv run on the folded view [B][L/r][r*C] with block-Toeplitz weights."""
    B, L, C, k, use_res = case
    torch.manual_seed(L * 7 + k)
    dev = torch.device("cuda:0")
    x = torch.randn(B, L, C, device=dev).bfloat16()
    w = torch.randn(C, C, k, device=dev) / (C * k) ** 0.5
    b = torch.randn(C, device=dev) * 0.1
    res = torch.randn(B, L, C, device=dev).bfloat16() if use_res else None
    y = ops.conv1d_cl(x, w, b, dilation=1, res=res, res_gain=10.0, out_slope=0.1, impl=0, desc_mode=16)
    y0 = ops.conv1d_cl(x, w, b, dilation=1, res=res, res_gain=10.0, out_slope=0.1, impl=0, desc_mode=0)
    torch.cuda.synchronize()
    assert rel_err(y, ref_conv(x, w, b, 1, res, 10.0, 0.1)) < REL_TOL
    # same bf16 operands, fp32 accumulation in a different order: at most one bf16 ulp apart
    assert float(((y.float() - y0.float()).abs() / (y0.float().abs() + 1e-3)).max()) < 2 ** -7


@pytest.mark.parametrize("case", [(2, 1000, 32, 3, 3, False), (3, 777, 32, 11, 5, True), (2, 3111, 32, 7, 3, False),
                                  (2, 517, 64, 7, 3, True), (1, 4000, 64, 11, 5, False), (1, 1, 32, 11, 5, False),
                                  (1, 7, 64, 3, 5, True), (16, 260, 32, 7, 5, False), (2, 2049, 64, 3, 3, False),
                                  (3, 3072, 32, 11, 3, True), (2, 2560, 64, 11, 5, True)],
                         ids=lambda c: "B%d_L%d_C%d_k%d_d%d_r%d" % c)
def test_dilated_time_folded_conv_matches_torch_and_unfolded(case):
    """Dilated convs fold per sub-sequence t = d*q + rho (5-d TMA view); ragged ends are masked in shared memory."""
    B, L, C, k, d, use_res = case
    torch.manual_seed(L * 7 + k + d)
    dev = torch.device("cuda:0")
    x = torch.randn(B, L, C, device=dev).bfloat16()
    w = torch.randn(C, C, k, device=dev) / (C * k) ** 0.5
    b = torch.randn(C, device=dev) * 0.1
    res = torch.randn(B, L, C, device=dev).bfloat16() if use_res else None
    y = ops.conv1d_cl(x, w, b, dilation=d, res=res, res_gain=10.0, out_slope=0.1, impl=0, desc_mode=16)
    y0 = ops.conv1d_cl(x, w, b, dilation=d, res=res, res_gain=10.0, out_slope=0.1, impl=0, desc_mode=0)
    torch.cuda.synchronize()
    assert torch.isfinite(y.float()).all()
    assert rel_err(y, ref_conv(x, w, b, d, res, 10.0, 0.1)) < REL_TOL
    assert float(((y.float() - y0.float()).abs() / (y0.float().abs() + 1e-3)).max()) < 2 ** -7


def ref_pair(x, w1, b1, w2, b2, d, slope=0.1):
    """One ResBlock1 iteration (modules.py:211-221) on the a-form input, h rounded to bf16 like the kernel stores it."""
    k = w1.shape[2]
    a = x.float().transpose(1, 2)
    h = F.conv1d(a, w1.bfloat16().float(), b1, dilation=d, padding=(k - 1) // 2 * d)
    h = torch.where(h >= 0, h, h * slope).bfloat16().float()
    y = F.conv1d(h, w2.bfloat16().float(), b2, padding=(k - 1) // 2)
    y = y + torch.where(a >= 0, a, a / slope)
    return torch.where(y >= 0, y, y * slope).transpose(1, 2)


@pytest.mark.parametrize("case", [(1, 300, 32, 3, 1), (2, 1000, 32, 7, 3), (3, 777, 32, 11, 5), (2, 250, 32, 11, 5),
                                  (1, 5, 32, 7, 1), (1, 246, 32, 11, 1), (1, 247, 32, 11, 3), (2, 3000, 64, 3, 3),
                                  (1, 254, 64, 3, 1), (1, 9000, 32, 3, 5), (2, 2000, 64, 7, 5), (1, 122, 64, 7, 1), (1, 123, 64, 7, 3)],
                         ids=lambda c: "B%d_L%d_C%d_k%d_d%d" % c)
def test_fused_resblock_pair_matches_torch_and_unfused(case):
    B, L, C, k, d = case
    torch.manual_seed(L + k)
    dev = torch.device("cuda:0")
    x = torch.randn(B, L, C, device=dev).bfloat16()
    w1 = torch.randn(C, C, k, device=dev) / (C * k) ** 0.5
    w2 = torch.randn(C, C, k, device=dev) / (C * k) ** 0.5
    b1 = torch.randn(C, device=dev) * 0.1
    b2 = torch.randn(C, device=dev) * 0.1
    y = ops.resblock_pair_cl(x, w1, b1, w2, b2, dilation=d, slope=0.1)
    torch.cuda.synchronize()
    assert rel_err(y, ref_pair(x, w1, b1, w2, b2, d)) < REL_TOL
    # identical arithmetic to the two single-conv launches (same bf16 h, same accumulation order)
    h = ops.conv1d_cl(x, w1, b1, dilation=d, out_slope=0.1)
    y2 = ops.conv1d_cl(h, w2, b2, dilation=1, res=x, res_gain=10.0, out_slope=0.1)
    assert torch.equal(y, y2)


@pytest.mark.parametrize("case", [(1, 1024, 32, 3, 1), (2, 4000, 32, 7, 1), (3, 3108, 32, 11, 1), (1, 8, 32, 7, 1),
                                  (2, 3000, 64, 3, 1), (2, 2000, 64, 7, 1), (1, 2222, 64, 11, 1), (16, 960, 32, 11, 1),
                                  (2, 3000, 64, 3, 3), (2, 2002, 64, 7, 3), (1, 4444, 64, 11, 3), (2, 2000, 64, 7, 2),
                                  # C = 128 on the plain view (r = 1): dilation = rows between taps of one tile
                                  (2, 1000, 128, 3, 1), (1, 2240, 128, 3, 3), (3, 225, 128, 3, 5), (1, 7, 128, 3, 1),
                                  (2, 900, 128, 7, 3), (1, 500, 128, 11, 1), (16, 448, 128, 3, 3), (2, 1500, 128, 11, 3),
                                  (2, 1100, 128, 7, 1), (3, 700, 128, 11, 5)],
                         ids=lambda c: "B%d_L%d_C%d_k%d_d%d" % c)
def test_time_folded_fused_pair_matches_torch(case):
    B, L, C, k, d = case
    torch.manual_seed(L + k + d)
    dev = torch.device("cuda:0")
    x = torch.randn(B, L, C, device=dev).bfloat16()
    w1 = torch.randn(C, C, k, device=dev) / (C * k) ** 0.5
    w2 = torch.randn(C, C, k, device=dev) / (C * k) ** 0.5
    b1 = torch.randn(C, device=dev) * 0.1
    b2 = torch.randn(C, device=dev) * 0.1
    y = ops.resblock_pair_cl(x, w1, b1, w2, b2, dilation=d, slope=0.1, folded=True)
    torch.cuda.synchronize()
    assert torch.isfinite(y.float()).all()
    assert rel_err(y, ref_pair(x, w1, b1, w2, b2, d)) < 1.5e-2   # h is re-rounded to bf16 from a differently ordered sum
    # against the two single-conv launches: same values up to the bf16 rounding of h and of the output
    h = ops.conv1d_cl(x, w1, b1, dilation=d, out_slope=0.1)
    y2 = ops.conv1d_cl(h, w2, b2, dilation=1, res=x, res_gain=10.0, out_slope=0.1)
    assert rel_err(y, y2) < 1.5e-2


@pytest.mark.parametrize("case", [
    # (B, L, [(k, dilation), ...]): one branch = one ResBlock1 iteration, three = the last pairs of the MRF + average
    (1, 512, [(3, 1)]), (2, 1000, [(7, 3)]), (3, 778, [(11, 5)]), (2, 500, [(11, 1)]), (1, 2, [(3, 1)]), (1, 6, [(7, 1)]),
    (1, 498, [(11, 3)]), (1, 502, [(11, 3)]), (1, 9000, [(3, 5)]), (16, 960, [(7, 1)]), (2, 3000, [(5, 2)]), (1, 4098, [(9, 4)]),
    (2, 1200, [(3, 5), (7, 5), (11, 5)]), (1, 256, [(3, 5), (7, 5), (11, 5)]), (3, 2, [(3, 1), (7, 3), (11, 5)]),
    (1, 5000, [(3, 1), (5, 3)]), (2, 2048, [(11, 5), (3, 5), (7, 5)]), (40, 600, [(3, 5), (7, 5), (11, 5)]),
    # C = 64 on plain rows (third field): one pair, weights resident up to k = 7
    (2, 1000, [(3, 1)], 64), (1, 777, [(7, 3)], 64), (3, 251, [(3, 5)], 64), (1, 3, [(7, 1)], 64), (16, 300, [(5, 2)], 64),
    (2, 2001, [(3, 1), (3, 3)], 64),
    # C = 128 (conv_mrf128.cu): streamed weights, channels-as-M tiles of 224-240 rows; the shipped stage-1 tail, permuted
    # branches, a single pair, lengths around the tile edge and shorter than the halo
    (2, 1000, [(3, 5), (7, 5), (11, 5)], 128), (3, 2240, [(11, 5), (3, 5), (7, 5)], 128), (1, 300, [(3, 1)], 128),
    (16, 448, [(7, 3)], 128), (1, 5, [(3, 1), (7, 1)], 128), (2, 225, [(11, 1), (11, 3)], 128), (1, 223, [(5, 2)], 128)],
    ids=lambda c: "B%d_L%d_%s%s" % (c[0], c[1], "+".join("k%dd%d" % kd for kd in c[2]), "_C%d" % c[3] if len(c) > 3 else ""))
def test_folded_narrow_stage_kernel_matches_torch(case):
    """conv_mrfp.cu (C = 32 on the 2-sample folded view): single pairs with dilation-1 (N = 64 chunk jobs) and dilated
    (N = 32 block jobs, odd and even dilations) first convs, and the three-branch form with the average; lengths around
    the 500-sample tile edge, shorter than one folded row's halo, and small problems that take the 256-sample tiles."""
    B, L, branches = case[:3]
    C = case[3] if len(case) > 3 else 32
    torch.manual_seed(L + 7 * len(branches))
    dev = torch.device("cuda:0")
    xs, w1s, b1s, w2s, b2s = [], [], [], [], []
    for k, d in branches:
        xs.append(torch.randn(B, L, C, device=dev).bfloat16())
        w1s.append(torch.randn(C, C, k, device=dev) / (C * k) ** 0.5)
        w2s.append(torch.randn(C, C, k, device=dev) / (C * k) ** 0.5)
        b1s.append(torch.randn(C, device=dev) * 0.1)
        b2s.append(torch.randn(C, device=dev) * 0.1)
    out_slope = 0.1 if len(branches) == 1 else 0.01
    y = ops.mrf_pairs_cl(xs, w1s, b1s, w2s, b2s, [d for _, d in branches], slope=0.1, out_slope=out_slope)
    torch.cuda.synchronize()
    assert torch.isfinite(y.float()).all()
    total = 0
    for (k, d), x, w1, b1, w2, b2 in zip(branches, xs, w1s, b1s, w2s, b2s):
        a = x.float().transpose(1, 2)
        h = F.conv1d(a, w1.bfloat16().float(), b1, dilation=d, padding=(k - 1) // 2 * d)
        h = torch.where(h >= 0, h, h * 0.1).bfloat16().float()
        total = total + F.conv1d(h, w2.bfloat16().float(), b2, padding=(k - 1) // 2) + torch.where(a >= 0, a, a / 0.1)
    total = total / len(branches)
    ref = torch.where(total >= 0, total, total * out_slope).transpose(1, 2)
    assert rel_err(y, ref) < 1.5e-2   # h is re-rounded to bf16 from a sum formed in another tap order
    if len(branches) == 1 and C <= 64:   # and against the plain-tile pair kernel of the same iteration
        k, d = branches[0]
        if k % 2 == 1 and k <= 15:
            y2 = ops.resblock_pair_cl(xs[0], w1s[0], b1s[0], w2s[0], b2s[0], dilation=d, slope=0.1)
            assert rel_err(y, y2) < 1.5e-2


def test_randomised_conv_shapes():
    """Seeded sweep over the supported shape space (channels, taps, dilation, ragged lengths, residual on/off, plain and
    time-folded forms): every case against torch on the same bf16 operands."""
    import numpy as np
    rs = np.random.RandomState(2024)
    dev = torch.device("cuda:0")
    n_fold = 0
    for case in range(48):
        ci = int(rs.choice([32, 64, 96, 128, 192, 256]))
        co = ci if rs.rand() < 0.7 else int(rs.choice([32, 64, 128, 256]))
        k = int(rs.choice([1, 3, 5, 7, 11]))
        d = int(rs.choice([1, 1, 2, 3, 5]))
        B = int(rs.randint(1, 4))
        L = int(rs.randint(1, 700))
        use_res = bool(rs.rand() < 0.5)
        fold = ci == co and ci in (32, 64) and rs.rand() < 0.6
        if fold and d == 1:
            L = max(4, L // 4 * 4)          # the dilation-1 folded form needs whole folded rows
        torch.manual_seed(case)
        x = torch.randn(B, L, ci, device=dev).bfloat16()
        w = torch.randn(co, ci, k, device=dev) / (ci * k) ** 0.5
        b = torch.randn(co, device=dev) * 0.1
        res = torch.randn(B, L, co, device=dev).bfloat16() if use_res else None
        y = ops.conv1d_cl(x, w, b, dilation=d, res=res, res_gain=10.0, out_slope=0.1, impl=0, desc_mode=16 if fold else 0)
        torch.cuda.synchronize()
        err = rel_err(y, ref_conv(x, w, b, d, res, 10.0, 0.1))
        assert torch.isfinite(y.float()).all() and err < REL_TOL, (case, B, L, ci, co, k, d, use_res, fold, err)
        n_fold += int(fold)
    assert n_fold >= 5


# ---- fp16 storage mode (decoder option "fp16"): same kernels, IEEE fp16 operands / residual / output
FP16_CASES = [
    (2, 300, 128, 128, 7, 5, True),     # channels-as-M tile, streamed weights, residual
    (2, 1000, 256, 256, 11, 5, True),   # two M-tiles of channels
    (3, 777, 32, 32, 11, 3, True),      # 64-byte rows, time-as-M
    (2, 517, 64, 64, 7, 3, True),
    (2, 50, 192, 512, 7, 1, False),     # conv_pre shape
    (1, 1, 64, 64, 11, 5, True),
    (5, 130, 96, 96, 3, 1, False),      # generic epilogue
]


@pytest.mark.parametrize("impl", [0, 1], ids=["tcgen05", "simt"])
@pytest.mark.parametrize("case", FP16_CASES, ids=lambda c: "B%d_L%d_ci%d_co%d_k%d_d%d_r%d" % c)
def test_conv1d_fp16_storage_matches_torch(case, impl):
    B, L, ci, co, k, d, use_res = case
    torch.manual_seed(B * 1000 + L + 7)
    dev = torch.device("cuda:0")
    x = torch.randn(B, L, ci, device=dev).half()
    w = torch.randn(co, ci, k, device=dev) / (ci * k) ** 0.5
    b = torch.randn(co, device=dev) * 0.1
    res = torch.randn(B, L, co, device=dev).half() if use_res else None
    y = ops.conv1d_cl(x, w, b, dilation=d, res=res, res_gain=10.0, out_slope=0.1, impl=impl)
    torch.cuda.synchronize()
    assert y.dtype == torch.float16
    kk = w.shape[2]
    r = F.conv1d(x.float().transpose(1, 2), w.half().float(), b, dilation=d, padding=(kk - 1) // 2 * d)
    if res is not None:
        rr = res.float().transpose(1, 2)
        r = r + torch.where(rr >= 0, rr, rr * 10.0)
    r = torch.where(r >= 0, r, r * 0.1).transpose(1, 2)
    # fp16 output rounding is 2^-12 relative: eight times tighter than the bf16 tolerance
    assert rel_err(y, r) < REL_TOL / 8


def test_fp16_storage_folded_forms_and_saturation():
    dev = torch.device("cuda:0")
    torch.manual_seed(11)
    # time-folded (plain and dilated) forms agree with the plain tile to one fp16 ulp
    for (L, k, d) in ((1024, 7, 1), (3111, 11, 3)):
        x = torch.randn(2, L, 32, device=dev).half()
        w = torch.randn(32, 32, k, device=dev) / (32 * k) ** 0.5
        b = torch.randn(32, device=dev) * 0.1
        res = torch.randn(2, L, 32, device=dev).half()
        a = ops.conv1d_cl(x, w, b, dilation=d, res=res, out_slope=0.1, impl=0).float()
        c = ops.conv1d_cl(x, w, b, dilation=d, res=res, out_slope=0.1, impl=0, desc_mode=16).float()
        assert float(((a - c).abs() / (a.abs() + 1e-2)).max()) < 2 ** -9
    # stores saturate at the largest finite fp16 instead of overflowing to inf
    x = torch.full((1, 256, 64), 200.0, device=dev).half()
    w = torch.full((64, 64, 3), 4.0, device=dev)
    y = ops.conv1d_cl(x, w, None, out_slope=0.1, impl=0)
    torch.cuda.synchronize()
    assert torch.isfinite(y.float()).all() and float(y.float().max()) == 65504.0
    y = ops.conv1d_cl(x, -w, None, out_slope=1.0, impl=0)
    assert torch.isfinite(y.float()).all() and float(y.float().min()) == -65504.0

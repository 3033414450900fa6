"""Python wrappers of the single-op C-ABI entries (per-kernel parity tests)."""
import torch

from . import _capi


def _ptr(t):
    return None if t is None else t.data_ptr()


def conv1d_cl(x, w, bias=None, dilation=1, res=None, res_gain=10.0, out_slope=1.0, impl=0, desc_mode=0):
    """Fused conv on channels-last bf16 activations.  x: bf16 [B, L, C_in]; w: fp32 [C_out, C_in, k] (rounded to
    bf16 inside); bias fp32 [C_out]; res: bf16 [B, L, C_out] stored post-leaky-relu(1/res_gain).  -> bf16 [B, L, C_out].
    A float16 x selects the fp16 storage mode (the decoder's "fp16" option): weights, res and y are fp16 as well."""
    assert x.is_cuda and x.dtype in (torch.bfloat16, torch.float16) and x.is_contiguous()
    assert res is None or res.dtype == x.dtype
    if x.dtype == torch.float16:
        desc_mode |= 1024
    B, L, c_in = x.shape
    c_out, c_in2, k = w.shape
    assert c_in2 == c_in
    w = w.float().contiguous()
    bias = None if bias is None else bias.float().contiguous()
    y = torch.empty((B, L, c_out), dtype=x.dtype, device=x.device)
    if (desc_mode & 16) and dilation > 1:
        # the dilated folded view reads (and masks) up to dilation*r rows past the end of x: give it NaN slack, like
        # the decoder's workspace does with its tail bytes
        buf = torch.full((x.numel() + 4096,), float("nan"), dtype=x.dtype, device=x.device)
        buf[:x.numel()].copy_(x.reshape(-1))
        x = buf[:x.numel()].view(B, L, c_in)
    st = torch.cuda.current_stream(x.device).cuda_stream
    L_ = _capi.lib(testing=bool(impl))   # impl=1 (CUDA-core cross-check) lives in the test build only
    _capi.check(L_.vitsdec_op_conv1d(x.device.index or 0, _ptr(x), _ptr(w), _ptr(bias), _ptr(res),
                                     float(res_gain), float(out_slope), _ptr(y), B, L, c_in, c_out, k,
                                     int(dilation), int(impl), int(desc_mode), st), "vitsdec_op_conv1d", L_)
    return y


def conv_transpose1d_cl(x, w, bias=None, stride=2, out_slope=1.0, impl=0):
    """Polyphase ConvTranspose1d(k, stride, padding=(k-stride)//2).  x: bf16 [B, L, C_in]; w: fp32 [C_in, C_out, k].
    -> bf16 [B, L*stride, C_out]."""
    assert x.is_cuda and x.dtype == torch.bfloat16 and x.is_contiguous()
    B, L, c_in = x.shape
    c_in2, c_out, k = w.shape
    assert c_in2 == c_in
    w = w.float().contiguous()
    bias = None if bias is None else bias.float().contiguous()
    y = torch.empty((B, L * stride, c_out), dtype=torch.bfloat16, device=x.device)
    st = torch.cuda.current_stream(x.device).cuda_stream
    L_ = _capi.lib(testing=bool(impl))
    _capi.check(L_.vitsdec_op_conv_transpose1d(x.device.index or 0, _ptr(x), _ptr(w), _ptr(bias),
                                               float(out_slope), _ptr(y), B, L, c_in, c_out, k, int(stride),
                                               int(impl), st), "vitsdec_op_conv_transpose1d", L_)
    return y


def resblock_pair_cl(x, w1, b1, w2, b2, dilation=1, slope=0.1, folded=False):
    """Fused ResBlock1 iteration.  x: bf16 [B, L, C] a-form; w1, w2: fp32 [C, C, k]; -> bf16 [B, L, C] a-form.
    folded=True runs the time-folded kernel (conv_pairf.cu)."""
    assert x.is_cuda and x.dtype == torch.bfloat16 and x.is_contiguous()
    B, L, C = x.shape
    k = w1.shape[2]
    y = torch.empty_like(x)
    if folded:
        if dilation > 1:  # the dilated view reads (and masks) a few rows past the end of x: NaN slack, see conv1d_cl
            buf = torch.full((x.numel() + 4096,), float("nan"), dtype=torch.bfloat16, device=x.device)
            buf[:x.numel()].copy_(x.reshape(-1))
            x = buf[:x.numel()].view(B, L, C)
        st = torch.cuda.current_stream(x.device).cuda_stream
        ts = [t.float().contiguous() for t in (w1, b1, w2, b2)]
        _capi.check(_capi.lib().vitsdec_op_resblock_pair_folded(x.device.index or 0, _ptr(x), _ptr(ts[0]), _ptr(ts[1]),
                                                                _ptr(ts[2]), _ptr(ts[3]), _ptr(y), B, L, C, k,
                                                                int(dilation), float(slope), st),
                    "vitsdec_op_resblock_pair_folded")
        return y
    st = torch.cuda.current_stream(x.device).cuda_stream
    ts = [t.float().contiguous() for t in (w1, b1, w2, b2)]
    _capi.check(_capi.lib().vitsdec_op_resblock_pair(x.device.index or 0, _ptr(x), _ptr(ts[0]), _ptr(ts[1]), _ptr(ts[2]),
                                                     _ptr(ts[3]), _ptr(y), B, L, C, k, int(dilation), float(slope), st),
                "vitsdec_op_resblock_pair")
    return y


def mrf_pairs_cl(xs, w1s, b1s, w2s, b2s, dilations, slope=0.1, out_slope=0.1):
    """conv_mrfp.cu on its own: len(xs) ResBlock1 iterations (C = 32, even length) summed and averaged.
    xs[j]: bf16 [B, L, 32] a-form; w1s[j], w2s[j]: fp32 [32, 32, k_j]; c1_j has dilations[j].  -> bf16 [B, L, 32]."""
    import ctypes
    nbr = len(xs)
    x0 = xs[0]
    assert all(x.is_cuda and x.dtype == torch.bfloat16 and x.is_contiguous() and x.shape == x0.shape for x in xs)
    B, L, C = x0.shape
    y = torch.empty_like(x0)
    keep = [[t.float().contiguous() for t in ts] for ts in (w1s, b1s, w2s, b2s)]

    def arr(ts):
        return (ctypes.c_void_p * nbr)(*[t.data_ptr() for t in ts])

    ks = (ctypes.c_int * nbr)(*[int(w.shape[2]) for w in w1s])
    ds = (ctypes.c_int * nbr)(*[int(d) for d in dilations])
    st = torch.cuda.current_stream(x0.device).cuda_stream
    _capi.check(_capi.lib().vitsdec_op_mrf_pairs(x0.device.index or 0, nbr, arr(xs), arr(keep[0]), arr(keep[1]), arr(keep[2]),
                                                 arr(keep[3]), _ptr(y), B, L, C, ks, ds, float(slope), float(out_slope), st),
                "vitsdec_op_mrf_pairs")
    return y

"""numpy restatement of the reference decoder forward.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/models.py:270-289 (Generator.forward),
/root/reference/modules.py:210-223 (ResBlock1.forward), :246-252 (ResBlock2.forward),
/root/reference/commons.py:14-15 (get_padding).  The tensor ops the reference delegates to
torch (conv1d, conv_transpose1d, leaky_relu, tanh, weight_norm) are restated from their
definitions; layout is the reference's NCL.  Works in any float dtype (fp64 gives the
"infinite precision" answer the tolerance in tests is quoted against).
"""
import numpy as np

from .hparams import DecoderHParams
from .weights import fold_state_dict

LRELU_SLOPE = 0.1  # modules.py:17


def get_padding(kernel_size, dilation=1):  # commons.py:14-15
    return int((kernel_size * dilation - dilation) / 2)


def leaky_relu(x, slope):
    return np.where(x >= 0, x, x * np.asarray(slope, dtype=x.dtype))


def conv1d(x, w, b=None, dilation=1, padding=0):
    """x [B,Ci,L], w [Co,Ci,k] -> [B,Co,L + 2p - d(k-1)] (stride 1, zero padding)."""
    B, Ci, L = x.shape
    Co, Ci2, k = w.shape
    assert Ci == Ci2
    xp = np.zeros((B, Ci, L + 2 * padding), dtype=x.dtype)
    xp[:, :, padding:padding + L] = x
    Lo = L + 2 * padding - dilation * (k - 1)
    y = np.zeros((B, Co, Lo), dtype=x.dtype)
    for j in range(k):
        y += np.matmul(w[None, :, :, j], xp[:, :, j * dilation:j * dilation + Lo])
    if b is not None:
        y += b[None, :, None]
    return y


def conv_transpose1d(x, w, b=None, stride=1, padding=0):
    """x [B,Ci,L], w [Ci,Co,k] -> [B,Co,(L-1)s - 2p + k]; y[s*i - p + j] += w[ci,co,j] x[ci,i]."""
    B, Ci, L = x.shape
    Ci2, Co, k = w.shape
    assert Ci == Ci2
    full = np.zeros((B, Co, (L - 1) * stride + k), dtype=x.dtype)
    for j in range(k):
        full[:, :, j:j + (L - 1) * stride + 1:stride] += np.matmul(w[:, :, j].T[None], x)
    y = full[:, :, padding:full.shape[2] - padding]
    if b is not None:
        y = y + b[None, :, None]
    return y


def resblock1(x, sd, prefix, k, dil):  # modules.py:210-223 with x_mask=None
    for m in range(3):
        xt = leaky_relu(x, LRELU_SLOPE)
        xt = conv1d(xt, sd[prefix + "convs1.%d.weight" % m], sd[prefix + "convs1.%d.bias" % m],
                    dilation=dil[m], padding=get_padding(k, dil[m]))
        xt = leaky_relu(xt, LRELU_SLOPE)
        xt = conv1d(xt, sd[prefix + "convs2.%d.weight" % m], sd[prefix + "convs2.%d.bias" % m],
                    dilation=1, padding=get_padding(k, 1))
        x = xt + x
    return x


def resblock2(x, sd, prefix, k, dil):  # modules.py:246-252 with x_mask=None
    for m in range(len(dil)):
        xt = leaky_relu(x, LRELU_SLOPE)
        xt = conv1d(xt, sd[prefix + "convs.%d.weight" % m], sd[prefix + "convs.%d.bias" % m],
                    dilation=dil[m], padding=get_padding(k, dil[m]))
        x = xt + x
    return x


def generator_forward_np(hp: DecoderHParams, sd, z, g=None, dtype=np.float64, taps=None):
    """Reference Generator.forward(x, g) (models.py:270-289).

    ``sd``: state_dict as numpy arrays, weight-norm (233-key) or folded (157-key) form.
    ``taps``: optional dict that receives named intermediates (for per-layer parity tests).
    """
    if any(k.endswith("weight_v") for k in sd):
        sd = fold_state_dict(sd)
    sd = {k: np.asarray(v).astype(dtype) for k, v in sd.items()}
    x = np.asarray(z).astype(dtype)
    x = conv1d(x, sd["conv_pre.weight"], sd["conv_pre.bias"], padding=3)            # :271
    if g is not None:                                                                 # :272-273
        x = x + conv1d(np.asarray(g).astype(dtype), sd["cond.weight"], sd["cond.bias"])
    if taps is not None:
        taps["conv_pre"] = x
    nk = len(hp.resblock_kernel_sizes)
    rb = resblock1 if hp.resblock == "1" else resblock2
    for i, (u, k) in enumerate(zip(hp.upsample_rates, hp.upsample_kernel_sizes)):    # :275
        x = leaky_relu(x, LRELU_SLOPE)                                                # :276
        x = conv_transpose1d(x, sd["ups.%d.weight" % i], sd["ups.%d.bias" % i],
                             stride=u, padding=(k - u) // 2)                          # :277
        if taps is not None:
            taps["ups.%d" % i] = x
        xs = None
        for j in range(nk):                                                           # :279-283
            y = rb(x, sd, "resblocks.%d." % (i * nk + j),
                   hp.resblock_kernel_sizes[j], hp.resblock_dilation_sizes[j])
            xs = y if xs is None else xs + y
        x = xs / np.asarray(nk, dtype=dtype)                                          # :284
        if taps is not None:
            taps["mrf.%d" % i] = x
    x = leaky_relu(x, 0.01)                                                           # :285 (default slope!)
    x = conv1d(x, sd["conv_post.weight"], None, padding=3)                            # :286
    return np.tanh(x)                                                                 # :287

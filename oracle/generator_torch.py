"""torch-functional restatement of the reference decoder.  TEST INFRASTRUCTURE ONLY.

The reference's CPU implementation of this path *is* PyTorch eager: its ~120 lines
(/root/reference/models.py:244-289, modules.py:187-229) only compose aten
``conv1d`` / ``conv_transpose1d`` / ``leaky_relu`` / ``tanh`` and the old-style
``weight_norm`` hook, which re-materialises every weight on every forward
(``torch._weight_norm``; 76 calls per forward).  This module
issues exactly that op sequence through ``torch.nn.functional`` from a plain
state_dict, so that

* ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` arm times the same oneDNN kernels the
  reference would run on the box's host cores (``/root/reference`` itself does not
  exist on the GPU box), and
* GPU parity tests have an fp32 reference at sizes where the numpy oracle is slow.

It is validated against the real reference module in tests/golden/make_golden.py
(bit-identical on CPU) and against the committed golden vectors in tests/test_oracle.py.
"""
import torch
import torch.nn.functional as F

from .hparams import DecoderHParams

LRELU_SLOPE = 0.1  # modules.py:17


def _pad(k, d=1):  # commons.py:14-15
    return int((k * d - d) / 2)


def _w(sd, prefix):
    """Effective weight of a (possibly weight-normed) conv, recomputed per call like the
    reference's forward pre-hook does (torch.nn.utils.weight_norm, dim=0)."""
    if prefix + ".weight_v" in sd:
        return torch._weight_norm(sd[prefix + ".weight_v"], sd[prefix + ".weight_g"], 0)
    return sd[prefix + ".weight"]


def _resblock1(x, sd, p, k, dil):  # modules.py:210-223, x_mask=None
    for m in range(3):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, _w(sd, p + "convs1.%d" % m), sd[p + "convs1.%d.bias" % m],
                      dilation=dil[m], padding=_pad(k, dil[m]))
        xt = F.leaky_relu(xt, LRELU_SLOPE)
        xt = F.conv1d(xt, _w(sd, p + "convs2.%d" % m), sd[p + "convs2.%d.bias" % m],
                      dilation=1, padding=_pad(k, 1))
        x = xt + x
    return x


def _resblock2(x, sd, p, k, dil):  # modules.py:246-252, x_mask=None
    for m in range(len(dil)):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, _w(sd, p + "convs.%d" % m), sd[p + "convs.%d.bias" % m],
                      dilation=dil[m], padding=_pad(k, dil[m]))
        x = xt + x
    return x


@torch.no_grad()
def generator_forward_torch(hp: DecoderHParams, sd, z, g=None, taps=None):
    """Generator.forward(x, g) (models.py:270-289) from a state_dict of torch tensors."""
    x = F.conv1d(z, sd["conv_pre.weight"], sd["conv_pre.bias"], padding=3)
    if g is not None:
        x = x + F.conv1d(g, sd["cond.weight"], sd["cond.bias"])
    if taps is not None:
        taps["conv_pre"] = x
    nk = len(hp.resblock_kernel_sizes)
    rb = _resblock1 if hp.resblock == "1" else _resblock2
    for i, (u, k) in enumerate(zip(hp.upsample_rates, hp.upsample_kernel_sizes)):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, _w(sd, "ups.%d" % i), sd["ups.%d.bias" % i],
                               stride=u, padding=(k - u) // 2)
        if taps is not None:
            taps["ups.%d" % i] = x
        xs = None
        for j in range(nk):
            y = rb(x, sd, "resblocks.%d." % (i * nk + j),
                   hp.resblock_kernel_sizes[j], hp.resblock_dilation_sizes[j])
            if xs is None:
                xs = y
            else:
                xs += y
        x = xs / nk
        if taps is not None:
            taps["mrf.%d" % i] = x
    x = F.leaky_relu(x)  # default slope 0.01, models.py:285
    x = F.conv1d(x, sd["conv_post.weight"], None, padding=3)
    return torch.tanh(x)


def to_torch_state_dict(sd_np, dtype=torch.float32, device="cpu"):
    return {k: torch.from_numpy(v.copy()).to(device=device, dtype=dtype) for k, v in sd_np.items()}

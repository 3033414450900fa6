"""Deterministic decoder weights in the reference's state_dict form.  TEST INFRASTRUCTURE ONLY.

Key names / shapes follow the parameter tree built at /root/reference/models.py:245-268
and /root/reference/modules.py:187-208 (ResBlock1), :232-243 (ResBlock2) with the
old-style ``torch.nn.utils.weight_norm`` (``weight_g`` / ``weight_v``; dim=0).

There is no network for checkpoints, so tests / bench synthesise weights with
``numpy.random.RandomState`` (bit-stable across numpy versions) instead of torch's
RNG.  Scales mimic torch's default Conv init (U(+-1/sqrt(fan_in))) and
``weight_g`` is *perturbed* away from ``||v||`` so that a fold which ignores
``g`` (or normalises over the wrong axis) fails parity.
"""
import numpy as np

from .hparams import DecoderHParams


def _stage_channels(hp: DecoderHParams):
    c0 = hp.upsample_initial_channel
    return [c0 // (2 ** (i + 1)) for i in range(len(hp.upsample_rates))]


def state_dict_keys(hp: DecoderHParams, weight_norm: bool = True):
    """(key, shape) list in the reference's registration order."""
    out = []
    c0 = hp.upsample_initial_channel
    out.append(("conv_pre.weight", (c0, hp.initial_channel, 7)))
    out.append(("conv_pre.bias", (c0,)))

    def wn(prefix, wshape):
        if weight_norm:
            out.append((prefix + ".bias", (wshape[1] if prefix.startswith("ups") else wshape[0],)))
            out.append((prefix + ".weight_g", (wshape[0], 1, 1)))
            out.append((prefix + ".weight_v", wshape))
        else:
            out.append((prefix + ".weight", wshape))
            out.append((prefix + ".bias", (wshape[1] if prefix.startswith("ups") else wshape[0],)))

    for i, (u, k) in enumerate(zip(hp.upsample_rates, hp.upsample_kernel_sizes)):
        wn("ups.%d" % i, (c0 // (2 ** i), c0 // (2 ** (i + 1)), k))  # ConvTranspose1d: [C_in, C_out, k]
    n = 0
    for ch in _stage_channels(hp):
        for k, dil in zip(hp.resblock_kernel_sizes, hp.resblock_dilation_sizes):
            if hp.resblock == "1":
                for m in range(3):
                    wn("resblocks.%d.convs1.%d" % (n, m), (ch, ch, k))
                for m in range(3):
                    wn("resblocks.%d.convs2.%d" % (n, m), (ch, ch, k))
            else:
                for m in range(len(dil)):
                    wn("resblocks.%d.convs.%d" % (n, m), (ch, ch, k))
            n += 1
    out.append(("conv_post.weight", (1, _stage_channels(hp)[-1], 7)))
    if hp.gin_channels:
        out.append(("cond.weight", (c0, hp.gin_channels, 1)))
        out.append(("cond.bias", (c0,)))
    return out


def synth_state_dict(hp: DecoderHParams, seed: int = 0, weight_norm: bool = True, gain: float = 1.0):
    """Seeded fp32 numpy state_dict with the reference's keys (233 tensors for the shipped config).

    ``gain`` scales the effective ResBlock conv weights so activations do not
    collapse towards zero through 26 layers of default-init convs.
    """
    rs = np.random.RandomState(seed)
    sd = {}
    for key, shape in state_dict_keys(hp, weight_norm=True):
        if key.endswith(".bias"):
            sd[key] = rs.uniform(-0.05, 0.05, size=shape).astype(np.float32)
        elif key.endswith("weight_g"):
            sd[key] = None  # filled after weight_v
        else:
            if key.startswith("ups"):
                fan_in = shape[1] * shape[2]  # torch uses size(1)*k for ConvTranspose
            else:
                fan_in = shape[1] * shape[2]
            bound = 1.0 / np.sqrt(fan_in)
            w = rs.uniform(-bound, bound, size=shape).astype(np.float32)
            if key.startswith("resblocks"):
                w *= np.float32(gain)
            sd[key] = w
            if key.endswith("weight_v"):
                gk = key[: -len("weight_v")] + "weight_g"
                norm = np.sqrt((w.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True))
                sd[gk] = (norm * rs.uniform(0.5, 1.5, size=norm.shape)).astype(np.float32)
    if not weight_norm:
        sd = fold_state_dict(sd)
    return sd


def fold_weight_norm(v: np.ndarray, g: np.ndarray) -> np.ndarray:
    """w = v * g / ||v||_2 over every dim but 0 (torch.nn.utils.weight_norm, dim=0).

    Conv1d: dim 0 is C_out.  ConvTranspose1d: dim 0 is C_in (weight is [C_in, C_out, k]) --
    /root/reference/models.py:254 applies weight_norm with the default dim=0 to both.
    """
    v64 = v.astype(np.float64)
    norm = np.sqrt((v64 ** 2).sum(axis=tuple(range(1, v.ndim)), keepdims=True))
    return (v64 * (g.astype(np.float64) / norm)).astype(v.dtype)


def fold_state_dict(sd):
    """233-key weight-norm form -> 157-key plain form (what remove_weight_norm() leaves,
    /root/reference/models.py:291-296)."""
    out = {}
    for k, v in sd.items():
        if k.endswith("weight_g"):
            continue
        if k.endswith("weight_v"):
            base = k[: -len("weight_v")]
            out[base + "weight"] = fold_weight_norm(v, sd[base + "weight_g"])
        else:
            out[k] = v
    return out

"""CPU restatement of the reference's normalising flow between the prior and the decoder.  TEST INFRASTRUCTURE ONLY.

``ResidualCouplingBlock.forward(x, x_mask, g, reverse)`` (/root/reference/models.py:179-209) is what produces the latent
``z`` that ``Generator`` decodes (models.py:521-522): ``n_flows`` x [ResidualCouplingLayer (modules.py:298-343,
mean_only=True), Flip (modules.py:270-277)], each coupling layer running a WaveNet-style ``WN`` (modules.py:111-184) with
the gate ``fused_add_tanh_sigmoid_multiply`` (commons.py:103-110).  The arithmetic itself lives in PyTorch (conv1d, tanh,
sigmoid, old-style weight_norm with dim=0); this file only composes the same ops functionally, weight norm recomputed per
call like the reference does.  Pinned against the unmodified reference by tests/golden/make_golden_flow.py.
"""
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class FlowHParams:
    """Constructor arguments of ResidualCouplingBlock as SynthesizerTrn passes them (models.py:449)."""
    channels: int = 192
    hidden_channels: int = 192
    kernel_size: int = 5
    dilation_rate: int = 1
    n_layers: int = 4
    n_flows: int = 4
    gin_channels: int = 256

    def ctor_args(self):
        return (self.channels, self.hidden_channels, self.kernel_size, self.dilation_rate, self.n_layers), \
            {"n_flows": self.n_flows, "gin_channels": self.gin_channels}


FLOW_FINETUNE_SPEAKER = FlowHParams()                       # models.py:449 with configs/finetune_speaker.json
FLOW_TINY = FlowHParams(64, 64, 3, 2, 2, 2, 32)             # small shapes for fast tests (not a reference config)
FLOW_TINY_NOG = FlowHParams(64, 32, 5, 1, 3, 3, 0)


def flow_state_dict_keys(hp: FlowHParams):
    """(key, shape) in the reference's registration order (modules.py:121-147, 318-322)."""
    out = []
    H, C2 = hp.hidden_channels, hp.channels // 2
    for i in range(hp.n_flows):
        p = "flows.%d." % (2 * i)      # odd entries are Flip modules without parameters (models.py:199-201)
        out.append((p + "pre.weight", (H, C2, 1)))
        out.append((p + "pre.bias", (H,)))
        for l in range(hp.n_layers):
            out.append((p + "enc.in_layers.%d.bias" % l, (2 * H,)))
            out.append((p + "enc.in_layers.%d.weight_g" % l, (2 * H, 1, 1)))
            out.append((p + "enc.in_layers.%d.weight_v" % l, (2 * H, H, hp.kernel_size)))
        for l in range(hp.n_layers):
            rs = 2 * H if l < hp.n_layers - 1 else H
            out.append((p + "enc.res_skip_layers.%d.bias" % l, (rs,)))
            out.append((p + "enc.res_skip_layers.%d.weight_g" % l, (rs, 1, 1)))
            out.append((p + "enc.res_skip_layers.%d.weight_v" % l, (rs, H, 1)))
        if hp.gin_channels:
            n = 2 * H * hp.n_layers
            out.append((p + "enc.cond_layer.bias", (n,)))
            out.append((p + "enc.cond_layer.weight_g", (n, 1, 1)))
            out.append((p + "enc.cond_layer.weight_v", (n, hp.gin_channels, 1)))
        out.append((p + "post.weight", (C2, H, 1)))
        out.append((p + "post.bias", (C2,)))
    return out


def synth_flow_state_dict(hp: FlowHParams, seed: int = 0):
    """Seeded fp32 numpy state_dict.  ``post`` is NOT zero (the reference zero-initialises it, modules.py:321-322, which
    would make every coupling layer an identity) and ``weight_g`` is perturbed away from ||v|| so the fold is exercised."""
    rs = np.random.RandomState(seed)
    sd = {}
    for key, shape in flow_state_dict_keys(hp):
        if key.endswith(".bias"):
            sd[key] = rs.uniform(-0.1, 0.1, size=shape).astype(np.float32)
        elif key.endswith("weight_g"):
            sd[key] = None
        else:
            bound = 1.0 / np.sqrt(shape[1] * shape[2])
            w = rs.uniform(-bound, bound, size=shape).astype(np.float32)
            sd[key] = w
            if key.endswith("weight_v"):
                norm = np.sqrt((w.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True))
                sd[key[: -len("weight_v")] + "weight_g"] = (norm * rs.uniform(0.5, 1.5, size=norm.shape)).astype(np.float32)
    return sd


def _wn(sd, prefix):
    v, g = sd[prefix + ".weight_v"], sd[prefix + ".weight_g"]
    return v * (g / v.flatten(1).norm(dim=1).view(-1, 1, 1))   # torch.nn.utils.weight_norm, dim=0


def _wn_forward(hp, sd, p, x, x_mask, g):
    """WN.forward, modules.py:149-176."""
    H = hp.hidden_channels
    output = torch.zeros_like(x)
    if g is not None:
        g = F.conv1d(g, _wn(sd, p + "cond_layer"), sd[p + "cond_layer.bias"])
    for i in range(hp.n_layers):
        dil = hp.dilation_rate ** i
        pad = int((hp.kernel_size * dil - dil) / 2)
        x_in = F.conv1d(x, _wn(sd, p + "in_layers.%d" % i), sd[p + "in_layers.%d.bias" % i], dilation=dil, padding=pad)
        g_l = g[:, i * 2 * H:(i + 1) * 2 * H, :] if g is not None else torch.zeros_like(x_in)
        in_act = x_in + g_l                                               # commons.py:105-110
        acts = torch.tanh(in_act[:, :H, :]) * torch.sigmoid(in_act[:, H:, :])
        rs = F.conv1d(acts, _wn(sd, p + "res_skip_layers.%d" % i), sd[p + "res_skip_layers.%d.bias" % i])
        if i < hp.n_layers - 1:
            x = (x + rs[:, :H, :]) * x_mask
            output = output + rs[:, H:, :]
        else:
            output = output + rs
    return output * x_mask


def flow_forward_torch(hp: FlowHParams, sd, x, x_mask, g=None, reverse=False):
    """ResidualCouplingBlock.forward (models.py:203-210).  sd: torch state_dict; x [B, C, T]; x_mask [B, 1, T]."""
    C2 = hp.channels // 2
    order = list(range(hp.n_flows))

    def coupling(i, x):  # ResidualCouplingLayer.forward with mean_only=True, modules.py:324-343
        p = "flows.%d." % (2 * i)
        x0, x1 = x[:, :C2], x[:, C2:]
        h = F.conv1d(x0, sd[p + "pre.weight"], sd[p + "pre.bias"]) * x_mask
        h = _wn_forward(hp, sd, p + "enc.", h, x_mask, g)
        m = F.conv1d(h, sd[p + "post.weight"], sd[p + "post.bias"]) * x_mask
        x1 = (m + x1 * x_mask) if not reverse else (x1 - m) * x_mask       # logs = 0
        return torch.cat([x0, x1], 1)

    if not reverse:
        for i in order:
            x = torch.flip(coupling(i, x), [1])
    else:
        for i in reversed(order):
            x = coupling(i, torch.flip(x, [1]))
    return x

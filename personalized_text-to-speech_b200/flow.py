"""Drop-in ``ResidualCouplingBlock`` (the flow that produces the decoder's latent), backed by libvitsdec.so.

Mirrors /root/reference/models.py:179-209: same constructor, same parameter tree / ``state_dict`` keys
(``flows.<2i>.pre``, ``flows.<2i>.enc.{in_layers,res_skip_layers}.<l>``, ``flows.<2i>.enc.cond_layer``,
``flows.<2i>.post``; odd entries are parameter-free ``Flip`` modules) and the same
``forward(x, x_mask, g=None, reverse=False)`` contract (forward returns ``x`` only, like the reference's block).
The torch sub-modules only HOLD parameters; the arithmetic runs in ``vitsdec_flow_apply`` (csrc/flow.cu).  There is
no PyTorch fallback; training (autograd) is outside the contract.
"""
import ctypes
import threading
import warnings

import torch
from torch import nn
from torch.nn.utils import weight_norm

from . import _capi


class _WNParams(nn.Module):
    """Parameter holder with WN's tree (modules.py:111-147)."""

    def __init__(self, hidden_channels, kernel_size, dilation_rate, n_layers, gin_channels=0):
        super().__init__()
        self.in_layers = nn.ModuleList()
        self.res_skip_layers = nn.ModuleList()
        if gin_channels != 0:
            self.cond_layer = weight_norm(nn.Conv1d(gin_channels, 2 * hidden_channels * n_layers, 1), name="weight")
        for i in range(n_layers):
            dilation = dilation_rate ** i
            padding = int((kernel_size * dilation - dilation) / 2)
            self.in_layers.append(weight_norm(nn.Conv1d(hidden_channels, 2 * hidden_channels, kernel_size,
                                                        dilation=dilation, padding=padding), name="weight"))
            rs = 2 * hidden_channels if i < n_layers - 1 else hidden_channels
            self.res_skip_layers.append(weight_norm(nn.Conv1d(hidden_channels, rs, 1), name="weight"))


class _CouplingParams(nn.Module):
    """Parameter holder with ResidualCouplingLayer's tree (modules.py:298-322, mean_only=True)."""

    def __init__(self, channels, hidden_channels, kernel_size, dilation_rate, n_layers, gin_channels=0):
        super().__init__()
        half = channels // 2
        self.pre = nn.Conv1d(half, hidden_channels, 1)
        self.enc = _WNParams(hidden_channels, kernel_size, dilation_rate, n_layers, gin_channels=gin_channels)
        self.post = nn.Conv1d(hidden_channels, half, 1)
        self.post.weight.data.zero_()   # modules.py:321-322
        self.post.bias.data.zero_()


class _Flip(nn.Module):
    """modules.py:270-277 (no parameters); the channel reversal itself happens inside the native block."""


class ResidualCouplingBlock(nn.Module):
    """``ResidualCouplingBlock(channels, hidden_channels, kernel_size, dilation_rate, n_layers, n_flows=4,
    gin_channels=0)`` -- models.py:180-201."""

    def __init__(self, channels, hidden_channels, kernel_size, dilation_rate, n_layers, n_flows=4, gin_channels=0):
        super().__init__()
        self.channels = int(channels)
        self.hidden_channels = int(hidden_channels)
        self.kernel_size = int(kernel_size)
        self.dilation_rate = int(dilation_rate)
        self.n_layers = int(n_layers)
        self.n_flows = int(n_flows)
        self.gin_channels = int(gin_channels)
        self.flows = nn.ModuleList()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")  # old-style weight_norm gives the reference's weight_g / weight_v keys
            for _ in range(n_flows):
                self.flows.append(_CouplingParams(channels, hidden_channels, kernel_size, dilation_rate, n_layers,
                                                  gin_channels=gin_channels))
                self.flows.append(_Flip())
        self._handle = None
        self._handle_device = None
        self._loaded_fingerprint = None
        self._lock = threading.RLock()   # held across weight sync AND the decode enqueue, see generator.py
        self.assume_frozen = False
        self._options = {}
        self._plist = None
        self._dirty = True
        self._ws = {}
        self._register_load_state_dict_pre_hook(self._mark_dirty)

    def set_option(self, key, value):
        """fp16: 1 = fp16 instead of bf16 conv operands / stored activations (the latent stays fp32 either way);
        the weights are re-folded on the next forward."""
        changed = self._options.get(key) != int(value)
        self._options[key] = int(value)
        if self._handle is not None:
            _capi.check(_capi.lib().vitsdec_flow_set_option(self._handle, key.encode(), int(value)), "flow set_option")
        if changed and key == "fp16":
            self._loaded_fingerprint = None

    def _mark_dirty(self, *args):
        self._dirty = True

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._plist = None
        self._dirty = True
        self._ws = {}
        return out

    def _fingerprint(self):
        """Sum of the parameters' version counters over a cached list (see Generator._fingerprint)."""
        if self._plist is None:
            self._plist = list(self.parameters())
        pl = self._plist
        return (len(pl), sum(p._version for p in pl), pl[0].data_ptr(), pl[-1].data_ptr())

    def _workspace(self, device, nbytes):
        key = torch.cuda.current_stream(device).cuda_stream
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes or ws.device != device:
            if ws is None and len(self._ws) >= 4:
                self._ws.pop(next(iter(self._ws)))
            # geometric growth, as in Generator._workspace: a new pointer invalidates the stream's cached plans
            grown = 0 if ws is None or ws.device != device else ws.numel() + ws.numel() // 2
            ws = None
            self._ws.pop(key, None)
            ws = torch.empty(max(nbytes, grown), dtype=torch.uint8, device=device)
            self._ws[key] = ws
        return ws

    def _sync_native(self, device):
        lib = _capi.lib()
        if self._handle is None or self._handle_device != device:
            if self._handle is not None:
                lib.vitsdec_flow_destroy(self._handle)
                self._handle = None
            hp = _capi.FlowHParams(self.channels, self.hidden_channels, self.kernel_size, self.dilation_rate,
                                   self.n_layers, self.n_flows, self.gin_channels)
            h = _capi._vp()
            index = device.index if device.index is not None else torch.cuda.current_device()
            _capi.check(lib.vitsdec_flow_create(ctypes.byref(hp), index, ctypes.byref(h)), "vitsdec_flow_create")
            self._handle, self._handle_device, self._loaded_fingerprint = h, device, None
            for k, v in self._options.items():
                _capi.check(lib.vitsdec_flow_set_option(self._handle, k.encode(), v), "flow set_option")
        fp = None
        if self._loaded_fingerprint is None or self._dirty or not self.assume_frozen:
            fp = self._fingerprint()
        if self._loaded_fingerprint is None or (fp is not None and fp != self._loaded_fingerprint):
            torch.cuda.synchronize(device)   # no decode in flight may still read the packed weights
            mods = dict(self.named_modules())
            stream = torch.cuda.current_stream(device).cuda_stream
            keep = []
            for i in range(lib.vitsdec_flow_num_layers(self._handle)):
                name = lib.vitsdec_flow_layer_name(self._handle, i).decode()
                m = mods[name]
                tensors = (m.weight_v, m.weight_g, m.bias) if hasattr(m, "weight_v") else (m.weight, None, m.bias)
                ptrs = []
                for t in tensors:
                    if t is None:
                        ptrs.append(None)
                        continue
                    if t.device != device:
                        raise RuntimeError("vitsdec flow: parameter %s is on %s but the input is on %s"
                                           % (name, t.device, device))
                    t = t.detach().float().contiguous()
                    keep.append(t)
                    ptrs.append(t.data_ptr())
                _capi.check(lib.vitsdec_flow_load_layer(self._handle, name.encode(), ptrs[0], ptrs[1], ptrs[2], stream),
                            "vitsdec_flow_load_layer(%s)" % name)
            torch.cuda.current_stream(device).synchronize()
            self._loaded_fingerprint = fp if fp is not None else self._fingerprint()
        self._dirty = False

    def __del__(self):
        try:
            if self._handle is not None:
                _capi.lib().vitsdec_flow_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    def forward(self, x, x_mask, g=None, reverse=False):
        """x: [B, channels, T] float on a CUDA device; x_mask: [B, 1, T] binary mask (commons.sequence_mask) or None;
        g: [B, gin_channels, 1] or None.  Returns [B, channels, T] (models.py:203-210)."""
        if x.dim() != 3 or x.shape[1] != self.channels:
            raise RuntimeError("ResidualCouplingBlock.forward: expected x of shape [B, %d, T], got %s"
                               % (self.channels, tuple(x.shape)))
        if not x.is_cuda:
            raise RuntimeError("ResidualCouplingBlock.forward: vitsdec has no CPU path (got x on %s)" % x.device)
        if torch.is_grad_enabled() and (x.requires_grad or self.training):
            raise RuntimeError("ResidualCouplingBlock.forward: vitsdec is inference-only (no autograd); call under "
                               "torch.no_grad() with the module in eval() mode")
        if g is not None and self.gin_channels == 0:
            raise RuntimeError("ResidualCouplingBlock.forward: g given but gin_channels=0")
        B, _, T = x.shape
        if B == 0 or T == 0:
            return x.clone()
        device = x.device
        out_dtype = x.dtype
        xf = x if x.dtype == torch.float32 else x.float()
        if xf.stride(2) != 1:
            xf = xf.contiguous()
        mk = None
        if x_mask is not None:
            if x_mask.numel() != B * T:
                raise RuntimeError("ResidualCouplingBlock.forward: expected x_mask of shape [%d, 1, %d], got %s"
                                   % (B, T, tuple(x_mask.shape)))
            mk = x_mask.to(device=device, dtype=torch.float32).reshape(B, T).contiguous()
        gf = None
        if g is not None:
            gf = g.to(device=device, dtype=torch.float32).reshape(B, self.gin_channels).contiguous()
        lib = _capi.lib()
        with torch.cuda.device(device), self._lock:
            self._sync_native(device)
            nbytes = lib.vitsdec_flow_workspace_bytes(self._handle, B, T)
            ws = self._workspace(device, nbytes)
            out = torch.empty((B, self.channels, T), dtype=torch.float32, device=device)
            stream = torch.cuda.current_stream(device).cuda_stream
            _capi.check(lib.vitsdec_flow_apply(self._handle, xf.data_ptr(), xf.stride(0), xf.stride(1),
                                               None if mk is None else mk.data_ptr(),
                                               None if gf is None else gf.data_ptr(), out.data_ptr(), B, T,
                                               1 if reverse else 0, ws.data_ptr(), nbytes, stream), "vitsdec_flow_apply")
        return out if out_dtype == torch.float32 else out.to(out_dtype)

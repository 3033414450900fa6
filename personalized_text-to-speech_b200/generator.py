"""Drop-in ``Generator`` for the reference's HiFi-GAN waveform decoder, backed by libvitsdec.so.

Mirrors /root/reference/models.py:244-296 (``Generator``) and modules.py:187-256 (``ResBlock1`` /
``ResBlock2``): same constructor signature, same parameter tree and ``state_dict`` keys (233 tensors
with weight norm for the shipped config, 157 after ``remove_weight_norm()``), same ``forward(x, g=None)``
contract.  The torch sub-modules below only HOLD parameters so that ``load_state_dict`` /
``utils.load_checkpoint`` (utils.py:148-180), ``.to()``, ``.eval()`` and optimizers see exactly the
reference's tensors; they are never called.  ``forward`` hands raw device pointers to the C ABI
(``vitsdec_decode``), which runs the hand-written sm_100a kernels.  There is no PyTorch fallback.
"""
import operator
import threading

import torch
from torch import nn
from torch.nn.utils import remove_weight_norm, weight_norm

from . import _capi

LRELU_SLOPE = 0.1  # modules.py:17
_VERSION = operator.attrgetter("_version")


def get_padding(kernel_size, dilation=1):  # commons.py:14-15
    return int((kernel_size * dilation - dilation) / 2)


class _ResBlock1Params(nn.Module):
    """Parameter holder with ResBlock1's tree (modules.py:187-208): convs1.{0,1,2}, convs2.{0,1,2}."""

    def __init__(self, channels, kernel_size=3, dilation=(1, 3, 5)):
        super().__init__()
        self.convs1 = nn.ModuleList([
            weight_norm(nn.Conv1d(channels, channels, kernel_size, 1, dilation=d,
                                  padding=get_padding(kernel_size, d))) for d in dilation[:3]])
        self.convs2 = nn.ModuleList([
            weight_norm(nn.Conv1d(channels, channels, kernel_size, 1, dilation=1,
                                  padding=get_padding(kernel_size, 1))) for _ in range(3)])

    def remove_weight_norm(self):
        for l in list(self.convs1) + list(self.convs2):
            remove_weight_norm(l)


class _ResBlock2Params(nn.Module):
    """Parameter holder with ResBlock2's tree (modules.py:232-243): convs.{0,1}."""

    def __init__(self, channels, kernel_size=3, dilation=(1, 3)):
        super().__init__()
        self.convs = nn.ModuleList([
            weight_norm(nn.Conv1d(channels, channels, kernel_size, 1, dilation=d,
                                  padding=get_padding(kernel_size, d))) for d in dilation])

    def remove_weight_norm(self):
        for l in self.convs:
            remove_weight_norm(l)


class Generator(nn.Module):
    """``Generator(initial_channel, resblock, resblock_kernel_sizes, resblock_dilation_sizes,
    upsample_rates, upsample_initial_channel, upsample_kernel_sizes, gin_channels=0)`` -- models.py:245."""

    def __init__(self, initial_channel, resblock, resblock_kernel_sizes, resblock_dilation_sizes, upsample_rates,
                 upsample_initial_channel, upsample_kernel_sizes, gin_channels=0):
        super().__init__()
        import warnings
        self.num_kernels = len(resblock_kernel_sizes)
        self.num_upsamples = len(upsample_rates)
        self._hp_args = (initial_channel, resblock, list(resblock_kernel_sizes),
                         [list(d) for d in resblock_dilation_sizes], list(upsample_rates),
                         upsample_initial_channel, list(upsample_kernel_sizes), gin_channels)
        self.initial_channel = int(initial_channel)
        self.gin_channels = int(gin_channels)
        self.hop = 1
        for u in upsample_rates:
            self.hop *= int(u)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")  # old-style weight_norm is what gives the reference's key names
            self.conv_pre = nn.Conv1d(initial_channel, upsample_initial_channel, 7, 1, padding=3)
            block = _ResBlock1Params if str(resblock) == "1" else _ResBlock2Params
            self.ups = nn.ModuleList()
            for i, (u, k) in enumerate(zip(upsample_rates, upsample_kernel_sizes)):
                self.ups.append(weight_norm(nn.ConvTranspose1d(
                    upsample_initial_channel // (2 ** i), upsample_initial_channel // (2 ** (i + 1)),
                    k, u, padding=(k - u) // 2)))
            self.resblocks = nn.ModuleList()
            ch = upsample_initial_channel
            for i in range(len(self.ups)):
                ch = upsample_initial_channel // (2 ** (i + 1))
                for k, d in zip(resblock_kernel_sizes, resblock_dilation_sizes):
                    self.resblocks.append(block(ch, k, d))
            self.conv_post = nn.Conv1d(ch, 1, 7, 1, padding=3, bias=False)
            if gin_channels != 0:
                self.cond = nn.Conv1d(gin_channels, upsample_initial_channel, 1)
        # accept the folded (remove_weight_norm) checkpoint form as well as the weight_g/weight_v form
        self._register_load_state_dict_pre_hook(self._accept_folded_keys)
        self._handle = None          # vitsdec_decoder*
        self._handle_device = None
        self._loaded_fingerprint = None
        # Gradio calls forward from worker threads (VC_inference.py:38-53).  The lock is held across the weight sync AND
        # the (asynchronous, ~0.1 ms) enqueue of the decode: a re-fold or a handle swap can then never run while another
        # thread is inside vitsdec_decode, and a re-fold first drains the device so that no decode still in flight on
        # another stream reads packed weights that are being rewritten.
        self._lock = threading.RLock()
        self.assume_frozen = False   # True: skip the per-call parameter-version check
        self._options = {}
        self._testing_lib = False    # option impl=1 (CUDA-core cross-check) switches this instance to libvitsdec_test.so
        self._plist = None           # cached parameter list (the module-tree walk of parameters() costs ~0.25 ms per call)
        self._dirty = True           # set by load_state_dict / .to() / remove_weight_norm: forces a re-fold check
        self._ws = {}                # CUDA stream -> workspace tensor kept across calls (stable pointer = plan / graph hits)

    # ------------------------------------------------------------------ state_dict compatibility
    def _accept_folded_keys(self, state_dict, prefix, *args):
        self._dirty = True
        own = dict(self.named_parameters())
        for name in list(own):
            if not name.endswith("weight_v"):
                continue
            base = prefix + name[: -len("weight_v")]
            if base + "weight" in state_dict and base + "weight_v" not in state_dict:
                w = state_dict.pop(base + "weight")
                state_dict[base + "weight_v"] = w
                state_dict[base + "weight_g"] = w.float().flatten(1).norm(dim=1).view(-1, 1, 1).to(w.dtype)

    def remove_weight_norm(self):  # models.py:291-296
        print('Removing weight norm...')
        for l in self.ups:
            remove_weight_norm(l)
        for l in self.resblocks:
            l.remove_weight_norm()
        self._loaded_fingerprint = None
        self._plist = None
        self._dirty = True

    def _apply(self, fn, *args, **kwargs):   # .to() / .cuda() / .float(): parameters move or are replaced
        out = super()._apply(fn, *args, **kwargs)
        self._plist = None
        self._dirty = True
        self._ws = {}
        return out

    # ------------------------------------------------------------------ native handle management
    def set_option(self, key, value):
        """impl: 0 = tcgen05 kernels (default), 1 = CUDA-core cross-check kernels; debug_keep: keep intermediates;
        fp16: 1 = fp16 instead of bf16 operands / stored activations (same speed, ~18 dB more waveform SNR, stores
        saturate at +-65504; the weights are re-folded on the next forward).  Full list: include/vitsdec.h."""
        changed = self._options.get(key) != int(value)
        self._options[key] = int(value)
        if key == "impl" and int(value) != 0 and not self._testing_lib:
            # the product library has no second backend: rebuild the handle on the test build (tests only)
            with self._lock:
                self._drop_handle()
                self._testing_lib = True
        if self._handle is not None:
            _capi.check(self._lib().vitsdec_set_option(self._handle, key.encode(), int(value)), "set_option", self._lib())
        if key == "fp16" and changed:
            self._loaded_fingerprint = None   # packed weights of the other 16-bit format are invalid

    def _lib(self):
        return _capi.lib(testing=self._testing_lib)

    def _drop_handle(self):
        if self._handle is not None:
            self._lib().vitsdec_destroy(self._handle)   # synchronises the device first
            self._handle = None
            self._loaded_fingerprint = None

    def _fingerprint(self):
        """Cheap change detector (~20 us for 233 tensors): every in-place update of a parameter bumps its version
        counter, so the sum over a fixed parameter list is strictly monotone; moves / replacements (which change
        pointers without touching versions) set ``_dirty`` through ``_apply`` / the state_dict hook instead."""
        if self._plist is None:
            self._plist = list(self.parameters())
        pl = self._plist
        return (len(pl), sum(map(_VERSION, pl)), pl[0].data_ptr(), pl[-1].data_ptr())

    def _layer_tensors(self):
        """state_dict prefix -> (weight or weight_v, weight_g or None, bias or None)."""
        out = {}
        mods = dict(self.named_modules())
        lib = self._lib()
        for i in range(lib.vitsdec_num_layers(self._handle)):
            name = lib.vitsdec_layer_name(self._handle, i).decode()
            m = mods[name]
            if hasattr(m, "weight_v"):
                out[name] = (m.weight_v, m.weight_g, m.bias)
            else:
                out[name] = (m.weight, None, m.bias)
        return out

    def _sync_native(self, device):
        """(Re)build the native decoder and (re)fold weights when parameters moved or changed."""
        lib = self._lib()
        if self._handle is None or self._handle_device != device:
            self._drop_handle()
            hp = _capi.make_hparams(*self._hp_args[:7], gin_channels=self._hp_args[7])
            h = _capi._vp()
            import ctypes
            index = device.index if device.index is not None else torch.cuda.current_device()
            _capi.check(lib.vitsdec_create(ctypes.byref(hp), index, ctypes.byref(h)), "vitsdec_create", lib)
            self._handle = h
            self._handle_device = device
            self._loaded_fingerprint = None
            for k, v in self._options.items():
                _capi.check(lib.vitsdec_set_option(self._handle, k.encode(), v), "set_option", lib)
        fp = None
        if self._loaded_fingerprint is None or self._dirty or not self.assume_frozen:
            fp = self._fingerprint()
        if self._loaded_fingerprint is None or (fp is not None and fp != self._loaded_fingerprint):
            # decodes of other threads / streams may still be reading the packed weights: drain the device first (rare:
            # once per load_state_dict / .to() / in-place update)
            torch.cuda.synchronize(device)
            stream = torch.cuda.current_stream(device).cuda_stream
            keep = []
            for name, (w, g, b) in self._layer_tensors().items():
                ts = []
                for t in (w, g, b):
                    if t is None:
                        ts.append(None)
                        continue
                    if t.device != device:
                        raise RuntimeError("vitsdec: parameter %s is on %s but the input is on %s" % (name, t.device, device))
                    t = t.detach().float().contiguous()
                    keep.append(t)
                    ts.append(t.data_ptr())
                _capi.check(lib.vitsdec_load_layer(self._handle, name.encode(), ts[0], ts[1], ts[2], stream),
                            "vitsdec_load_layer(%s)" % name, lib)
            torch.cuda.current_stream(device).synchronize()  # temporaries in `keep` die here
            self._loaded_fingerprint = fp if fp is not None else self._fingerprint()
        self._dirty = False

    def _workspace(self, device, nbytes):
        """Per-stream scratch kept across calls.  The C side caches its launch plans (tensor maps, CUDA graphs) per
        workspace pointer: a fresh torch.empty per call keeps the pointer only as long as the caching allocator happens
        to return the same block.  Grow-only, at most 4 streams are remembered; release_workspace() frees them."""
        key = torch.cuda.current_stream(device).cuda_stream
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes or ws.device != device:
            if ws is None and len(self._ws) >= 4:
                self._ws.pop(next(iter(self._ws)))
            # grow geometrically: a new workspace pointer invalidates every cached plan of this stream (the C side keys
            # plans on it), so a serving loop whose requests creep upwards must not re-allocate for every new maximum
            grown = 0 if ws is None or ws.device != device else ws.numel() + ws.numel() // 2
            ws = None                       # release the old block before asking for the larger one
            self._ws.pop(key, None)
            ws = torch.empty(max(nbytes, grown), dtype=torch.uint8, device=device)
            self._ws[key] = ws
        return ws

    def release_workspace(self):
        with self._lock:
            self._ws = {}

    def __del__(self):
        try:
            self._drop_handle()
        except Exception:
            pass

    # ------------------------------------------------------------------ forward
    def forward(self, x, g=None):
        """x: [B, initial_channel, T] float on a CUDA device; g: [B, gin_channels, 1] or None.
        Returns [B, 1, T * prod(upsample_rates)] (models.py:270-289)."""
        if x.dim() != 3 or x.shape[1] != self.initial_channel:
            raise RuntimeError("Generator.forward: expected x of shape [B, %d, T], got %s"
                               % (self.initial_channel, tuple(x.shape)))
        if not x.is_cuda:
            raise RuntimeError("Generator.forward: vitsdec has no CPU path; move the module and inputs to a B200 "
                               "(got x on %s)" % x.device)
        if torch.is_grad_enabled() and (x.requires_grad or self.training):
            raise RuntimeError("Generator.forward: vitsdec is inference-only (no autograd); call under "
                               "torch.no_grad() with the module in eval() mode")
        if g is not None and self.gin_channels == 0:
            raise RuntimeError("Generator.forward: g given but gin_channels=0")
        B, _, T = x.shape
        out_dtype = x.dtype
        if B == 0 or T == 0:
            return x.new_zeros((B, 1, T * self.hop))
        device = x.device
        xf = x if x.dtype == torch.float32 else x.float()
        if xf.stride(2) != 1:
            xf = xf.contiguous()
        gf = None
        if g is not None:
            if g.shape[0] != B or g.shape[1] != self.gin_channels:
                raise RuntimeError("Generator.forward: expected g of shape [%d, %d, 1], got %s"
                                   % (B, self.gin_channels, tuple(g.shape)))
            gf = g.to(device=device, dtype=torch.float32).reshape(B, self.gin_channels).contiguous()
        lib = self._lib()
        with torch.cuda.device(device), self._lock:
            self._sync_native(device)
            nbytes = lib.vitsdec_workspace_bytes(self._handle, B, T)
            ws = self._workspace(device, nbytes)
            out = torch.empty((B, 1, T * self.hop), dtype=torch.float32, device=device)
            stream = torch.cuda.current_stream(device).cuda_stream
            _capi.check(lib.vitsdec_decode(self._handle, xf.data_ptr(), xf.stride(0), xf.stride(1),
                                           None if gf is None else gf.data_ptr(), out.data_ptr(), B, T,
                                           ws.data_ptr(), nbytes, stream), "vitsdec_decode", lib)
        return out if out_dtype == torch.float32 else out.to(out_dtype)

    def receptive_halo(self):
        """Latent frames each side that an output sample depends on (chunked.receptive_halo; 13 for the shipped config)."""
        from .chunked import receptive_halo
        a = self._hp_args
        return receptive_halo(a[1], a[2], a[3], a[4], a[6])

    # ------------------------------------------------------------------ helpers for tests / bench
    @classmethod
    def from_reference(cls, ref_generator, *ctor_args, **ctor_kwargs):
        """Build from constructor args and copy a reference Generator's state_dict (post-hoc swap of net_g.dec)."""
        new = cls(*ctor_args, **ctor_kwargs)
        new.load_state_dict(ref_generator.state_dict())
        p = next(ref_generator.parameters())
        return new.to(p.device).train(ref_generator.training)

    def profile_read(self):
        """(accumulated conv-kernel device ms, conv launches) since set_option('profile', 1)."""
        import ctypes
        ms, n = ctypes.c_double(0), ctypes.c_int64(0)
        _capi.check(self._lib().vitsdec_profile_read(self._handle, ctypes.byref(ms), ctypes.byref(n)), "profile_read")
        return ms.value, n.value

    def last_launch_count(self):
        return 0 if self._handle is None else self._lib().vitsdec_last_launch_count(self._handle)

    def get_option(self, key):
        """Read an option / counter of the native decoder (include/vitsdec.h), e.g. "graph_failed": number of launch
        plans whose CUDA-graph capture failed and that therefore run as plain launches."""
        import ctypes
        if self._handle is None:
            return self._options.get(key, 0)
        v = ctypes.c_int(0)
        _capi.check(self._lib().vitsdec_get_option(self._handle, key.encode(), ctypes.byref(v)), "get_option", self._lib())
        return v.value

    def debug_read(self, name, batch, frames):
        """fp32 [B, C, L] copy of a kept intermediate of the last forward (set_option('debug_keep', 1)).
        Names: conv_pre, ups.<i>, mrf.<i>."""
        import ctypes
        lib = self._lib()
        c0, rates = self._hp_args[5], self._hp_args[4]
        if name == "conv_pre":
            C, L = c0, frames
        else:
            i = int(name.split(".")[1])
            C, L = c0 // (2 ** (i + 1)), frames
            for u in rates[: i + 1]:
                L *= u
        C = -(-C // 32) * 32   # stages narrower than 32 channels are carried zero-padded
        dev = self._handle_device
        buf = torch.empty((batch, C, L), dtype=torch.float32, device=dev)
        c, l = ctypes.c_int(0), ctypes.c_int(0)
        _capi.check(lib.vitsdec_debug_read(self._handle, name.encode(), buf.data_ptr(), buf.numel(), ctypes.byref(c),
                                           ctypes.byref(l), torch.cuda.current_stream(dev).cuda_stream), "debug_read")
        assert (c.value, l.value) == (C, L)
        torch.cuda.current_stream(dev).synchronize()
        return buf

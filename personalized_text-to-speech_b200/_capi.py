"""ctypes binding of libvitsdec.so (include/vitsdec.h).  Raw pointers and sizes only.

The library is mandatory: there is no Python / PyTorch fallback for the decode path.  If the shared
object is missing it is built with nvcc (personalized_text-to-speech_b200/build.py); if that is
impossible the import fails loudly.
"""
import ctypes
import os
import threading

from . import build as _build

MAX_UPSAMPLES = 8
MAX_KERNELS = 8
MAX_DILATIONS = 8


class HParams(ctypes.Structure):
    """struct vitsdec_hparams"""
    _fields_ = [
        ("initial_channel", ctypes.c_int32),
        ("resblock", ctypes.c_int32),
        ("num_kernels", ctypes.c_int32),
        ("resblock_kernel_sizes", ctypes.c_int32 * MAX_KERNELS),
        ("num_dilations", ctypes.c_int32 * MAX_KERNELS),
        ("resblock_dilation_sizes", (ctypes.c_int32 * MAX_DILATIONS) * MAX_KERNELS),
        ("num_upsamples", ctypes.c_int32),
        ("upsample_rates", ctypes.c_int32 * MAX_UPSAMPLES),
        ("upsample_initial_channel", ctypes.c_int32),
        ("upsample_kernel_sizes", ctypes.c_int32 * MAX_UPSAMPLES),
        ("gin_channels", ctypes.c_int32),
    ]


class FlowHParams(ctypes.Structure):
    """struct vitsdec_flow_hparams"""
    _fields_ = [(n, ctypes.c_int32) for n in ("channels", "hidden_channels", "kernel_size", "dilation_rate", "n_layers",
                                               "n_flows", "gin_channels")]


class VitsdecError(RuntimeError):
    pass


_lib = None
_lib_test = None
_lock = threading.Lock()

_vp, _i, _f, _sz, _i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_int64
_cp = ctypes.c_char_p

# name -> (restype, argtypes); must list every symbol include/vitsdec.h declares (tests check this)
SIGNATURES = {
    "vitsdec_abi_version": (_i, []),
    "vitsdec_last_error": (_cp, []),
    "vitsdec_create": (_i, [ctypes.POINTER(HParams), _i, ctypes.POINTER(_vp)]),
    "vitsdec_destroy": (None, [_vp]),
    "vitsdec_num_layers": (_i, [_vp]),
    "vitsdec_layer_name": (_cp, [_vp, _i]),
    "vitsdec_load_layer": (_i, [_vp, _cp, _vp, _vp, _vp, _vp]),
    "vitsdec_workspace_bytes": (_sz, [_vp, _i, _i]),
    "vitsdec_decode": (_i, [_vp, _vp, _i64, _i64, _vp, _vp, _i, _i, _vp, _sz, _vp]),
    "vitsdec_decode_host": (_i, [_vp, _vp, _vp, _vp, _i, _i]),
    "vitsdec_wav_pcm16": (_i, [_i, _vp, _vp, _i64, _vp]),
    "vitsdec_set_option": (_i, [_vp, _cp, _i]),
    "vitsdec_get_option": (_i, [_vp, _cp, ctypes.POINTER(_i)]),
    "vitsdec_last_launch_count": (_i, [_vp]),
    "vitsdec_profile_read": (_i, [_vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_i64)]),
    "vitsdec_debug_read": (_i, [_vp, _cp, _vp, _sz, ctypes.POINTER(_i), ctypes.POINTER(_i), _vp]),
    "vitsdec_debug_set_trace": (_i, [_vp]),
    "vitsdec_flow_create": (_i, [ctypes.POINTER(FlowHParams), _i, ctypes.POINTER(_vp)]),
    "vitsdec_flow_destroy": (None, [_vp]),
    "vitsdec_flow_num_layers": (_i, [_vp]),
    "vitsdec_flow_layer_name": (_cp, [_vp, _i]),
    "vitsdec_flow_load_layer": (_i, [_vp, _cp, _vp, _vp, _vp, _vp]),
    "vitsdec_flow_set_option": (_i, [_vp, _cp, _i]),
    "vitsdec_flow_workspace_bytes": (_sz, [_vp, _i, _i]),
    "vitsdec_flow_apply": (_i, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "vitsdec_op_conv1d": (_i, [_i, _vp, _vp, _vp, _vp, _f, _f, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "vitsdec_op_resblock_pair": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp]),
    "vitsdec_op_resblock_pair_folded": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp]),
    "vitsdec_op_mrf_pairs": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _f, _f, _vp]),
    "vitsdec_op_conv_transpose1d": (_i, [_i, _vp, _vp, _vp, _f, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
}


def library_path():
    return _build.LIB


def lib(testing=False):
    """Load (building first if needed) libvitsdec.so and declare the prototypes.

    testing=True loads libvitsdec_test.so instead: the same code plus the CUDA-core cross-check backend (option
    impl=1).  Only the GPU tests ask for it (Generator.set_option("impl", 1), ops.*(impl=1)); the product library has
    no second backend."""
    global _lib, _lib_test
    if testing:
        if _lib_test is not None:
            return _lib_test
        lib()  # builds / freshness check
        with _lock:
            if _lib_test is None:
                L = ctypes.CDLL(_build.LIB_TEST)
                _declare(L)
                _lib_test = L
        return _lib_test
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = os.environ.get("VITSDEC_LIB") or _build.LIB   # VITSDEC_LIB: A/B two builds on one box (tools/ab.sh)
        if not os.environ.get("VITSDEC_LIB") and not _build.is_fresh():
            # missing, or older than csrc/ + include/ (source digest in csrc/.build_stamp): rebuild rather than load a
            # stale binary silently (*.so is git-ignored but ships in-tree)
            try:
                _build.build()
            except Exception as e:
                if not os.path.exists(path):  # no nvcc and no prebuilt library: nothing to run the decoder with
                    raise ImportError(
                        "vitsdec: %s is missing and could not be built (%s). There is no CPU/PyTorch "
                        "fallback for the decoder." % (path, e))
                import warnings
                warnings.warn("vitsdec: %s is older than its sources and could not be rebuilt (%s); loading it anyway"
                              % (path, e))
        L = ctypes.CDLL(path)
        _declare(L)
        _lib = L
    return _lib


def _declare(L):
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError here = ABI drift between header and library
        fn.restype = res
        fn.argtypes = args
    if L.vitsdec_abi_version() != 1:
        raise ImportError("vitsdec: ABI version mismatch")


def check(rc, what="", L=None):
    if rc != 0:
        msg = (L or lib()).vitsdec_last_error()
        raise VitsdecError("%s: %s" % (what or "vitsdec", msg.decode() if msg else "unknown error"))


def make_hparams(initial_channel, resblock, resblock_kernel_sizes, resblock_dilation_sizes, upsample_rates,
                 upsample_initial_channel, upsample_kernel_sizes, gin_channels=0):
    """Generator.__init__ arguments (reference models.py:245) -> struct vitsdec_hparams."""
    hp = HParams()
    hp.initial_channel = int(initial_channel)
    hp.resblock = 1 if str(resblock) == "1" else 2  # models.py:250: anything but '1' selects ResBlock2
    ks = [int(k) for k in resblock_kernel_sizes]
    ds = [[int(x) for x in d] for d in resblock_dilation_sizes]
    if len(ks) > MAX_KERNELS or len(upsample_rates) > MAX_UPSAMPLES or any(len(d) > MAX_DILATIONS for d in ds):
        raise ValueError("vitsdec: hyper-parameter lists too long")
    if len(ks) != len(ds) or len(upsample_rates) != len(upsample_kernel_sizes):
        raise ValueError("vitsdec: hyper-parameter list lengths disagree")
    hp.num_kernels = len(ks)
    for j, k in enumerate(ks):
        hp.resblock_kernel_sizes[j] = k
        hp.num_dilations[j] = len(ds[j])
        for m, d in enumerate(ds[j]):
            hp.resblock_dilation_sizes[j][m] = d
    hp.num_upsamples = len(upsample_rates)
    for i, (u, k) in enumerate(zip(upsample_rates, upsample_kernel_sizes)):
        hp.upsample_rates[i] = int(u)
        hp.upsample_kernel_sizes[i] = int(k)
    hp.upsample_initial_channel = int(upsample_initial_channel)
    hp.gin_channels = int(gin_channels)
    return hp

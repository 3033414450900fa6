"""CPU: the C-ABI library loads, exports every symbol include/vitsdec.h declares, and fails loudly without a GPU."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import pytest
import torch

import vitsdec
from vitsdec import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vitsdec.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"VITSDEC_API\s+[\w\s\*]+?\b(vitsdec_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _capi.lib()
    names = declared_symbols()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), "libvitsdec.so does not export %s" % n
    # and the ctypes prototypes cover the whole header (no silently unbound entry point)
    assert sorted(_capi.SIGNATURES) == names
    assert lib.vitsdec_abi_version() == 1
    # the test build (CUDA-core cross-check backend linked in) exports the same ABI; the product library carries no
    # second backend: its SASS holds no conv_simt kernel
    tlib = _capi.lib(testing=True)
    for n in names:
        assert hasattr(tlib, n)
    from importlib import import_module
    b = import_module("personalized_text-to-speech_b200.build")
    syms = subprocess.run(["nm", "-C", b.LIB], stdout=subprocess.PIPE, text=True).stdout
    syms_t = subprocess.run(["nm", "-C", b.LIB_TEST], stdout=subprocess.PIPE, text=True).stdout
    assert "conv_simt" not in syms and "conv_simt" in syms_t


def test_hparams_struct_layout_matches_header():
    prog = '#include <stdio.h>\n#include "vitsdec.h"\nint main(){printf("%zu %zu %zu", sizeof(vitsdec_hparams),' \
           ' __builtin_offsetof(vitsdec_hparams, num_upsamples), __builtin_offsetof(vitsdec_hparams, gin_channels));}'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(prog)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        size, off_nu, off_gin = map(int, subprocess.check_output([exe]).split())
    assert size == ctypes.sizeof(_capi.HParams)
    assert off_nu == _capi.HParams.num_upsamples.offset
    assert off_gin == _capi.HParams.gin_channels.offset


def test_header_is_plain_c():
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write('#include "vitsdec.h"\nint main(void){return VITSDEC_ABI_VERSION - 1;}\n')
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), c, "-o",
                               os.path.join(d, "t")])


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    lib = _capi.lib()
    hp = _capi.make_hparams(192, "1", [3, 7, 11], [[1, 3, 5]] * 3, [8, 8, 2, 2], 512, [16, 16, 4, 4], gin_channels=256)
    h = ctypes.c_void_p()
    rc = lib.vitsdec_create(ctypes.byref(hp), 0, ctypes.byref(h))
    assert rc != 0 and not h.value
    assert lib.vitsdec_last_error()  # a message, not a silent fallback


def test_make_hparams_rejects_bad_lists():
    with pytest.raises(ValueError):
        _capi.make_hparams(192, "1", [3, 7], [[1, 3, 5]] * 3, [8, 8], 512, [16, 16])
    hp = _capi.make_hparams(192, "2", [3, 7], [[1, 3], [1, 2]], [8, 8], 512, [16, 16])
    assert hp.resblock == 2 and hp.num_kernels == 2 and hp.num_dilations[1] == 2
    assert hp.resblock_dilation_sizes[1][1] == 2 and hp.gin_channels == 0


def test_no_oracle_import_in_product():
    """The product package must never route through the CPU oracle (it would void the parity claim)."""
    pkg = os.path.join(ROOT, "personalized_text-to-speech_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


def test_library_sass_is_tcgen05_and_tma():
    """The built library's SASS (no GPU needed) carries the Blackwell mnemonics the design claims: tcgen05.mma (UTCHMMA)
    with mbarrier commits (UTCBAR), TMEM loads (LDTM), TMEM allocation (UTCATOMSWS), TMA tensor loads (UTMALDG) and the
    256-bit global stores of the pair kernel -- and no legacy mma.sync (HMMA) or wgmma path."""
    import shutil
    import subprocess
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    _capi.lib()   # builds the library when this test runs first
    sass = subprocess.run([exe, "-sass", _capi.library_path()], stdout=subprocess.PIPE, text=True, check=True).stdout
    for mnemonic in ("UTCHMMA", "UTCBAR", "LDTM", "UTCATOMSWS", "UTMALDG", "STG.E.ENL2.256"):
        assert mnemonic in sass, mnemonic
    assert " HMMA." not in sass and "WGMMA" not in sass and "HGMMA" not in sass
    kernels = [l for l in sass.splitlines() if "Function :" in l]
    assert sum("conv_tc_kernel" in k for k in kernels) >= 40 and any("conv_pair_kernel" in k for k in kernels)

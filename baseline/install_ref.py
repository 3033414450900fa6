#!/usr/bin/env python
"""Install the UNMODIFIED reference into git-ignored ``baseline/_ref/``.

The reference (MedivhJin01/Personalized_Text-to-Speech) is plain Python without a setup.py / pyproject, so
``pip install --target baseline/_ref /root/reference`` has nothing to build; this script does what that install
would: it copies the handful of files the decoder path and its caller need, byte for byte, and builds the one
native piece (``monotonic_align/core.pyx``, Cython + gcc, the way README.md:7-12 of the reference describes) so that
``import models`` works.  ``baseline/_ref/`` is listed in .gitignore (never committed: no reference source enters
the history) but NOT in .gpurunignore, so it travels to the GPU box with the snapshot like the built .so files.

Users:  tests/test_gpu_reference_infer.py (SynthesizerTrn.infer / voice_conversion / load_checkpoint through the
        patched classes on the B200), bench.py --impl reference (kind "_ref") and bench.py's cuDNN-eager keys.
Run:    python baseline/install_ref.py            (build container only: needs /root/reference)
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("VITSDEC_REFERENCE", "/root/reference")
FILES = ["models.py", "models_infer.py", "modules.py", "commons.py", "attentions.py", "transforms.py", "utils.py",
         "configs/finetune_speaker.json", "configs/uma_trilingual.json", "configs/modified_finetune_speaker.json",
         "monotonic_align/__init__.py", "monotonic_align/core.pyx", "monotonic_align/setup.py"]


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def installed():
    return all(os.path.exists(os.path.join(DST, f)) for f in FILES)


def install(verbose=True):
    if not os.path.isdir(SRC):
        raise RuntimeError("reference tree %s not found (it only exists in the build container)" % SRC)
    manifest = []
    for f in FILES:
        d = os.path.join(DST, f)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, f), d)
        manifest.append("%s  %s" % (sha(d), f))
    # README.md:7-12: cd monotonic_align; mkdir monotonic_align; python setup.py build_ext --inplace
    ma = os.path.join(DST, "monotonic_align")
    os.makedirs(os.path.join(ma, "monotonic_align"), exist_ok=True)
    so = [f for f in os.listdir(os.path.join(ma, "monotonic_align")) if f.startswith("core") and f.endswith(".so")]
    if not so:
        r = subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=ma, stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout)
            raise RuntimeError("building monotonic_align failed")
    with open(os.path.join(DST, "MANIFEST.sha256"), "w") as fh:
        fh.write("\n".join(manifest) + "\n")
    if verbose:
        print("reference installed into %s (%d files, sha256 in MANIFEST.sha256)" % (DST, len(FILES)))
    return DST


if __name__ == "__main__":
    install()

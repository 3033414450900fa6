"""Locate the unmodified reference installed under baseline/_ref (baseline/install_ref.py) for tests and bench.py.

The GPU box has no /root/reference: the install travels with the repo snapshot (git-ignored, not gpurun-ignored).
In the build container the install is refreshed from /root/reference when it is missing."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def ref_dir():
    """Path of the installed reference, or None when it is neither installed nor installable here."""
    if not os.path.exists(os.path.join(REF, "models.py")) and os.path.isdir("/root/reference"):
        sys.path.insert(0, os.path.join(ROOT, "baseline"))
        try:
            import install_ref
            install_ref.install(verbose=False)
        except Exception:
            return None
        finally:
            sys.path.remove(os.path.join(ROOT, "baseline"))
    return REF if os.path.exists(os.path.join(REF, "models.py")) else None


_REF_MODULES = ("models", "models_infer", "modules", "commons", "attentions", "transforms", "utils", "monotonic_align")


def import_reference(name="models"):
    """Import a module of the installed reference (fresh: the reference uses flat top-level module names)."""
    d = ref_dir()
    if d is None:
        raise ImportError("reference not installed: run python baseline/install_ref.py in the build container")
    if d not in sys.path:
        sys.path.insert(0, d)
    return importlib.import_module(name)


def forget_reference():
    for n in list(sys.modules):
        if n.split(".")[0] in _REF_MODULES and getattr(sys.modules[n], "__file__", "") and \
                str(getattr(sys.modules[n], "__file__", "")).startswith(REF):
            del sys.modules[n]
    if REF in sys.path:
        sys.path.remove(REF)

"""Make the reference's own scripts use the B200 decoder without editing them.

``SynthesizerTrn.__init__`` resolves ``Generator`` as a module global at construction time
(/root/reference/models.py:447, models_infer.py:355), so assigning ``models.Generator`` before the model is
built is enough for ``cmd_inference.py`` / ``VC_inference.py`` / ``SynthesizerTrn.infer`` to run unchanged.
"""
import importlib
import sys

from .flow import ResidualCouplingBlock
from .generator import Generator

_saved = {}
_saved_flow = {}


def patch_reference(module_names=("models", "models_infer"), flow=False):
    """Replace ``<module>.Generator`` in every importable reference module.  Returns the patched names.

    ``flow=True`` also replaces ``<module>.ResidualCouplingBlock`` (resolved the same way at models.py:449), so that
    ``SynthesizerTrn.infer`` runs flow(reverse) -> dec natively.  Inference only: training needs autograd through it."""
    done = []
    for name in module_names:
        mod = sys.modules.get(name)
        if mod is None:
            try:
                mod = importlib.import_module(name)
            except Exception:
                continue
        if flow and getattr(mod, "ResidualCouplingBlock", None) is not ResidualCouplingBlock:
            _saved_flow[name] = getattr(mod, "ResidualCouplingBlock", None)
            mod.ResidualCouplingBlock = ResidualCouplingBlock
        if getattr(mod, "Generator", None) is Generator:
            done.append(name)
            continue
        _saved[name] = getattr(mod, "Generator", None)
        mod.Generator = Generator
        done.append(name)
    return done


def unpatch_reference():
    for name, orig in list(_saved.items()):
        mod = sys.modules.get(name)
        if mod is not None and orig is not None:
            mod.Generator = orig
        del _saved[name]
    for name, orig in list(_saved_flow.items()):
        mod = sys.modules.get(name)
        if mod is not None and orig is not None:
            mod.ResidualCouplingBlock = orig
        del _saved_flow[name]

"""``python -m vitsdec.run <reference_script.py> [args...]`` -- run a reference entry point
(cmd_inference.py, VC_inference.py) byte-for-byte unchanged with ``models.Generator`` replaced.

The reference directory (where the script lives) is put on sys.path exactly as running the script
directly would do.  ``VITSDEC_FLOW=1`` in the environment also replaces ``models.ResidualCouplingBlock`` (the flow that
produces the decoder's latent), for inference scripts only: the training scripts need autograd through it.
"""
import os
import runpy
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print(__doc__)
        return 2
    script = os.path.abspath(argv[0])
    sys.path.insert(0, os.path.dirname(script))
    from .patch import patch_reference
    patched = patch_reference(flow=os.environ.get("VITSDEC_FLOW", "0") == "1")
    if not patched:
        raise SystemExit("vitsdec.run: could not import the reference's models module from %s" % os.path.dirname(script))
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())

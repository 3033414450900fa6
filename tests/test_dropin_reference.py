"""CPU, build container only (needs /root/reference; skipped elsewhere): the drop-in claim at the level of the
reference's own model class.  With ``vitsdec.patch_reference(flow=True)`` the UNMODIFIED ``models_infer.SynthesizerTrn``
builds its ``dec`` and ``flow`` from the B200 classes, exposes exactly the same ``state_dict`` keys and shapes as the
stock model (so ``utils.load_checkpoint`` finds every tensor, utils.py:155-177), and strict-loads the stock model's
weights.  Running ``infer`` needs a B200 and is covered by the GPU parity tests of the two modules."""
import json
import os
import sys

import pytest
import torch

import vitsdec

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "models_infer.py")),
                                reason="the reference tree only exists in the build container")


def _build(models_infer):
    cfg = json.load(open(os.path.join(REF, "configs", "finetune_speaker.json")))
    torch.manual_seed(0)
    return models_infer.SynthesizerTrn(68, cfg["data"]["filter_length"] // 2 + 1,
                                       cfg["train"]["segment_size"] // cfg["data"]["hop_length"],
                                       n_speakers=cfg["data"]["n_speakers"], **cfg["model"])


def test_synthesizer_builds_on_the_b200_classes_with_the_same_checkpoint_layout():
    sys.path.insert(0, REF)
    try:
        import models_infer
        stock = _build(models_infer)
        assert type(stock.dec).__module__ == "models_infer"
        done = vitsdec.patch_reference(("models_infer",), flow=True)
        assert done == ["models_infer"]
        try:
            ours = _build(models_infer)
        finally:
            vitsdec.unpatch_reference()
        assert isinstance(ours.dec, vitsdec.Generator) and isinstance(ours.flow, vitsdec.ResidualCouplingBlock)
        assert models_infer.Generator is not vitsdec.Generator   # unpatched again
        want = {k: tuple(v.shape) for k, v in stock.state_dict().items()}
        got = {k: tuple(v.shape) for k, v in ours.state_dict().items()}
        assert list(got) == list(want) and got == want
        ours.load_state_dict(stock.state_dict(), strict=True)
        n_dec = sum(1 for k in want if k.startswith("dec."))
        n_flow = sum(1 for k in want if k.startswith("flow."))
        assert (n_dec, n_flow) == (233, 124)
    finally:
        sys.path.remove(REF)
        sys.modules.pop("models_infer", None)

"""CPU oracle for the VITS waveform decoder (HiFi-GAN Generator) hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package (``personalized_text-to-speech_b200`` / ``vitsdec``).  The only callers
allowed are ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs.

What is restated here (reference = /root/reference, MedivhJin01/Personalized_Text-to-Speech):

* ``models.py:244-289``   Generator.__init__ / forward (conv_pre, cond, ups, MRF, conv_post, tanh)
* ``modules.py:17``       LRELU_SLOPE = 0.1
* ``modules.py:187-229``  ResBlock1 (x_mask is always None on this path)
* ``modules.py:232-256``  ResBlock2
* ``commons.py:14-15``    get_padding
* third-party arithmetic the reference only *composes* (torch==2.1.2 pinned in
  requirements.txt:8; not vendored under /root/reference): ``conv1d``,
  ``conv_transpose1d``, ``leaky_relu``, ``tanh`` and ``weight_norm(dim=0)``
  restated from their published definitions.

Parity pinning: the reference ships no tests / golden vectors.
The oracle is pinned against the *reference itself executed in the build
container*: ``tests/golden/make_golden.py`` imports the unmodified
``/root/reference/models_infer.py`` Generator, runs it on seeded inputs and commits
the outputs under ``tests/golden/``; ``tests/test_oracle.py`` checks both oracle
implementations against those fixtures.
"""
from .hparams import (DecoderHParams, FINETUNE_SPEAKER, UMA_TRILINGUAL, TINY, TINY_RB2, HIFIGAN_V2,  # noqa: F401
                      HIFIGAN_V3)
from .weights import synth_state_dict, state_dict_keys, fold_weight_norm  # noqa: F401
from .generator_np import generator_forward_np  # noqa: F401

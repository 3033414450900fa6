"""B200-native VITS waveform decoder (HiFi-GAN Generator) -- drop-in for
MedivhJin01/Personalized_Text-to-Speech ``models.Generator`` (reference models.py:244-296).

The directory name follows the build contract (``personalized_text-to-speech_b200``); because of the
hyphens import it as ``import vitsdec`` (alias package at the repo root) or with
``importlib.import_module("personalized_text-to-speech_b200")``.
"""
from . import _capi  # noqa: F401
from .generator import Generator  # noqa: F401
from .flow import ResidualCouplingBlock  # noqa: F401
from .patch import patch_reference, unpatch_reference  # noqa: F401
from .sharding import shard_range, decode_sharded, PeerGather  # noqa: F401
from .chunked import decode_chunked  # noqa: F401
from .pipeline import HostPipeline  # noqa: F401
from .wavout import WavBatchWriter, wav_header, synthesize_to_wav  # noqa: F401
from .build import build  # noqa: F401
from .hparams import generator_args, generator_args_from_config  # noqa: F401

__all__ = ["Generator", "ResidualCouplingBlock", "patch_reference", "unpatch_reference", "shard_range", "decode_sharded", "PeerGather",
           "decode_chunked", "HostPipeline", "WavBatchWriter", "wav_header", "synthesize_to_wav", "build", "generator_args", "generator_args_from_config"]

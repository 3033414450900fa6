"""``vitsdec`` -- importable alias of the ``personalized_text-to-speech_b200`` package."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("personalized_text-to-speech_b200")

Generator = _pkg.Generator
ResidualCouplingBlock = _pkg.ResidualCouplingBlock
patch_reference = _pkg.patch_reference
unpatch_reference = _pkg.unpatch_reference
shard_range = _pkg.shard_range
decode_sharded = _pkg.decode_sharded
PeerGather = _pkg.PeerGather
decode_chunked = _pkg.decode_chunked
HostPipeline = _pkg.HostPipeline
WavBatchWriter = _pkg.WavBatchWriter
wav_header = _pkg.wav_header
synthesize_to_wav = _pkg.synthesize_to_wav
build = _pkg.build
generator_args = _pkg.generator_args
generator_args_from_config = _pkg.generator_args_from_config
_capi = _pkg._capi
__all__ = list(_pkg.__all__)

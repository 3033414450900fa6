"""Decoder hyper-parameters: the reference's ``configs/*.json`` ``model`` block -> ``Generator`` constructor args.

Mirrors how the reference feeds the decoder: ``utils.get_hparams_from_file`` (utils.py:361-367) parses the JSON,
``**hps.model`` is splatted into ``SynthesizerTrn`` (cmd_inference.py:93-98) which passes seven of those values to
``Generator`` positionally plus ``gin_channels`` (models.py:447).
"""
import json

# configs/finetune_speaker.json:35-52 == configs/uma_trilingual.json:35-52 (decoder-relevant keys)
SHIPPED_MODEL_BLOCK = {
    "inter_channels": 192,
    "resblock": "1",
    "resblock_kernel_sizes": [3, 7, 11],
    "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    "upsample_rates": [8, 8, 2, 2],
    "upsample_initial_channel": 512,
    "upsample_kernel_sizes": [16, 16, 4, 4],
    "gin_channels": 256,
}


def generator_args(model_block=None):
    """(args, kwargs) for ``Generator(*args, **kwargs)`` exactly as models.py:447 builds ``self.dec``."""
    m = dict(SHIPPED_MODEL_BLOCK if model_block is None else model_block)
    args = (m["inter_channels"], m["resblock"], m["resblock_kernel_sizes"], m["resblock_dilation_sizes"],
            m["upsample_rates"], m["upsample_initial_channel"], m["upsample_kernel_sizes"])
    return args, {"gin_channels": m.get("gin_channels", 0)}


def generator_args_from_config(path):
    """Read a reference config file (configs/finetune_speaker.json, configs/uma_trilingual.json, ...)."""
    with open(path, "r", encoding="utf-8") as f:
        cfg = json.load(f)
    return generator_args(cfg["model"])

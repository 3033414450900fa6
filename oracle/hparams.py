"""Decoder hyper-parameters as the reference passes them to ``Generator.__init__``.

Follows the positional signature at /root/reference/models.py:245 and the
``model`` block of /root/reference/configs/finetune_speaker.json:35-52
(identical in configs/uma_trilingual.json:35-52).  TEST INFRASTRUCTURE ONLY.
"""
from dataclasses import dataclass, field
from typing import List


@dataclass(frozen=True)
class DecoderHParams:
    initial_channel: int
    resblock: str
    resblock_kernel_sizes: tuple
    resblock_dilation_sizes: tuple
    upsample_rates: tuple
    upsample_initial_channel: int
    upsample_kernel_sizes: tuple
    gin_channels: int = 0

    def ctor_args(self):
        """Positional args + kwargs exactly as models.py:447 passes them."""
        return (
            self.initial_channel,
            self.resblock,
            [int(k) for k in self.resblock_kernel_sizes],
            [list(d) for d in self.resblock_dilation_sizes],
            [int(u) for u in self.upsample_rates],
            self.upsample_initial_channel,
            [int(k) for k in self.upsample_kernel_sizes],
        ), {"gin_channels": self.gin_channels}

    @property
    def hop(self):
        h = 1
        for u in self.upsample_rates:
            h *= u
        return h


# configs/finetune_speaker.json:43-51 with inter_channels=192 (models.py:447 passes inter_channels first)
FINETUNE_SPEAKER = DecoderHParams(
    192, "1", (3, 7, 11), ((1, 3, 5), (1, 3, 5), (1, 3, 5)), (8, 8, 2, 2), 512, (16, 16, 4, 4), 256
)
# configs/uma_trilingual.json:35-52 -- the model block is identical
UMA_TRILINGUAL = FINETUNE_SPEAKER

# small shapes for fast CPU tests / committed golden vectors (not reference configs)
TINY = DecoderHParams(64, "1", (3, 5), ((1, 3, 5), (1, 2, 3)), (4, 2), 128, (8, 4), 32)
TINY_RB2 = DecoderHParams(64, "2", (3, 7), ((1, 3), (1, 2)), (4, 2), 128, (8, 4), 0)

# Other HiFi-GAN shapes through the same constructor: the published V2 / V3 generator configurations
# (jik876/hifi-gan config_v2.json / config_v3.json: mel input of 80 channels, no speaker conditioning)
HIFIGAN_V2 = DecoderHParams(80, "1", (3, 7, 11), ((1, 3, 5), (1, 3, 5), (1, 3, 5)), (8, 8, 2, 2), 128, (16, 16, 4, 4), 0)
HIFIGAN_V3 = DecoderHParams(80, "2", (3, 5, 7), ((1, 2), (2, 6), (3, 12)), (8, 8, 4), 256, (16, 16, 8), 0)

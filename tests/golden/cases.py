"""Golden-vector case list shared by the generator script (make_golden.py, build container only) and the tests
(CPU and GPU boxes; must not import /root/reference)."""
import numpy as np

import oracle

CASES = [
    # name, hparams, seed, B, T, use_g
    ("tiny_b2_t9", oracle.TINY, 11, 2, 9, True),
    ("tiny_b1_t1", oracle.TINY, 12, 1, 1, True),
    ("tiny_rb2_b2_t13", oracle.TINY_RB2, 13, 2, 13, False),
    ("full_b2_t32", oracle.FINETUNE_SPEAKER, 21, 2, 32, True),
    ("full_b1_t7_nog", oracle.FINETUNE_SPEAKER, 22, 1, 7, False),
    # other HiFi-GAN shapes through the reference's constructor (V2: stages down to 8 channels; V3: ResBlock2, 3 stages)
    ("v2like_b1_t6", oracle.HIFIGAN_V2, 23, 1, 6, False),
    ("v3like_b2_t5", oracle.HIFIGAN_V3, 24, 2, 5, False),
]


def weight_checksum(sd):
    acc = 0.0
    for k in sorted(sd):
        v = sd[k].astype(np.float64).ravel()
        acc += float((v * np.cos(np.arange(v.size) % 97)).sum())
    return np.float64(acc)



"""Golden-vector cases of the flow (ResidualCouplingBlock); shared by make_golden_flow.py and the tests."""
from oracle.flow_torch import FLOW_FINETUNE_SPEAKER, FLOW_TINY, FLOW_TINY_NOG

FLOW_CASES = [
    # name, hparams, seed, B, T, lengths (None = full), reverse
    ("flow_tiny_b2_t11_rev", FLOW_TINY, 51, 2, 11, (11, 7), True),
    ("flow_tiny_b1_t5_fwd", FLOW_TINY, 52, 1, 5, None, False),
    ("flow_tinynog_b3_t9_rev", FLOW_TINY_NOG, 53, 3, 9, (9, 1, 4), True),
    ("flow_full_b2_t40_rev", FLOW_FINETUNE_SPEAKER, 54, 2, 40, (40, 23), True),
    ("flow_full_b1_t17_fwd", FLOW_FINETUNE_SPEAKER, 55, 1, 17, None, False),
]

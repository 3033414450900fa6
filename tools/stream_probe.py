"""Streaming ceiling of conv_tc_kernel: k=1 (pointwise) convs at the stage shapes, to be run under
`ncu --metrics gpu__time_duration.sum` (MMA work is negligible, so the time is the memory/epilogue path)."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitsdec  # noqa: E402,F401

ops = importlib.import_module("personalized_text-to-speech_b200.ops")
dev = torch.device("cuda:0")
for (C, L) in ((32, 220672), (64, 110336), (128, 55168), (256, 6896)):
    x = torch.randn(16, L, C, device=dev).bfloat16()
    r = torch.randn(16, L, C, device=dev).bfloat16()
    for k in (1, 3):
        w = torch.randn(C, C, k, device=dev) / (C * k) ** 0.5
        b = torch.zeros(C, device=dev)
        for res, dm in ((None, 0), (r, 0), (r, 4)):
            y = ops.conv1d_cl(x, w, b, dilation=1, res=res, out_slope=0.1, impl=0, desc_mode=dm)
    torch.cuda.synchronize()
print("ok")

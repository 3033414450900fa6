"""Per-tile timeline of CTA 0 of conv_pair_kernel (needs a VITSDEC_TRACE=1 build)."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitsdec  # noqa: E402

ops = importlib.import_module("personalized_text-to-speech_b200.ops")
lib = vitsdec._capi.lib()
dev = torch.device("cuda:0")
trace = torch.zeros(256 * 12, dtype=torch.int64, device=dev)
for (C, L, k, d) in ((32, 220672, 3, 1), (32, 220672, 7, 3), (32, 220672, 11, 5), (64, 110336, 3, 1), (64, 110336, 7, 3)):
    x = torch.randn(16, L, C, device=dev).bfloat16()
    w1 = torch.randn(C, C, k, device=dev) / (C * k) ** 0.5
    w2 = torch.randn(C, C, k, device=dev) / (C * k) ** 0.5
    b = torch.zeros(C, device=dev)
    ops.resblock_pair_cl(x, w1, b, w2, b, dilation=d)
    trace.zero_()
    lib.vitsdec_debug_set_trace(trace.data_ptr())
    ops.resblock_pair_cl(x, w1, b, w2, b, dilation=d)
    lib.vitsdec_debug_set_trace(None)
    torch.cuda.synchronize()
    t = trace.view(256, 12).cpu()
    n = int((t[:, 3] > 0).sum())
    t = t[:n].double()
    base = t[0, 0]
    print("C=%d k=%d d=%d tiles %d" % (C, k, d, n))
    for i in range(n // 2, min(n, n // 2 + 4)):
        print("  tile %3d: c1 %7d-%7d  c2 %7d-%7d | epi1 %7d-%7d epi2 %7d-%7d | epi2 wait from %7d  epi1 wait from %7d  TMA issued %7d  c1 wait from %7d"
              % ((i,) + tuple(int(v - base) for v in t[i][:12])))
    s = slice(10, n - 5)
    per = (t[n - 5, 7] - t[10, 7]) / (n - 15)
    print("  cycles/tile %.0f | c1 issue %.0f  c2 issue %.0f | epi1 %.0f  epi2 %.0f | c1 end->epi1 start %.0f  epi1 end->c2 start %.0f  c2 end->epi2 start %.0f"
          % (per, (t[s, 1] - t[s, 0]).mean(), (t[s, 3] - t[s, 2]).mean(), (t[s, 5] - t[s, 4]).mean(), (t[s, 7] - t[s, 6]).mean(),
             (t[s, 4] - t[s, 1]).mean(), (t[s, 2] - t[s, 5]).mean(), (t[s, 6] - t[s, 3]).mean()))

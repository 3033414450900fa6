"""Per-tile timeline of CTA 0 of conv_pairf_kernel (needs a VITSDEC_TRACE=1 build)."""
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
for (C, L, k, d) in ((64, 110336, 3, 1), (64, 110336, 3, 3), (64, 110336, 7, 1), (64, 110336, 7, 3), (64, 110336, 11, 1),
                     (64, 110336, 11, 3), (128, 55168, 3, 1), (128, 55168, 3, 3), (32, 220672, 11, 1)):
    x = torch.randn(16, L, C, device=dev).bfloat16()
    w1 = torch.randn(C, C, k, device=dev) / (C * k) ** 0.5
    w2 = torch.randn(C, C, k, device=dev) / (C * k) ** 0.5
    b = torch.zeros(C, device=dev)
    ops.resblock_pair_cl(x, w1, b, w2, b, dilation=d, folded=True)
    trace.zero_()
    lib.vitsdec_debug_set_trace(trace.data_ptr())
    ops.resblock_pair_cl(x, w1, b, w2, b, dilation=d, folded=True)
    lib.vitsdec_debug_set_trace(None)
    torch.cuda.synchronize()
    t = trace.view(256, 12).cpu()
    n = int((t[:, 3] > 0).sum())
    t = t[:n].double()
    base = t[0, 8]
    print("C=%d k=%d d=%d tiles %d" % (C, k, d, n))
    for i in range(n // 2, min(n, n // 2 + 3)):
        r = [int(v - base) for v in t[i][:10]]
        print("  tile %3d: xwait %7d c1 %7d-%7d | hwait %7d c2 %7d-%7d | hepi %7d-%7d oepi %7d-%7d"
              % (i, r[8], r[0], r[1], r[9], r[2], r[3], r[4], r[5], r[6], r[7]))
    s = slice(5, n - 3)
    per = (t[n - 3, 7] - t[5, 7]) / (n - 8)
    m = lambda a, b_: float((t[s, a] - t[s, b_]).mean())
    print("  cycles/tile %.0f | x wait %.0f  c1 issue %.0f  c1 end->hepi start %.0f  hepi %.0f  hepi end->c2 start %.0f  c2 issue %.0f"
          "  c2 end->oepi start %.0f  oepi %.0f" % (per, m(0, 8), m(1, 0), m(4, 1), m(5, 4), m(2, 5), m(3, 2), m(6, 3), m(7, 6)))

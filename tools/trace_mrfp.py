"""Per-(tile, branch) timeline of CTA 0 of conv_mrfp_kernel (needs a VITSDEC_TRACE=1 build: VITSDEC_TRACE=1 python -m
personalized_text-to-speech_b200.build --force).  Stamps: c1 / c2 issue, epi1 (h producer) / epi2 (output) work, waits."""
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
C, L = 32, 220672
for branches in ([(3, 1)], [(3, 3)], [(7, 1)], [(7, 3)], [(11, 1)], [(11, 3)], [(3, 5), (7, 5), (11, 5)]):
    nbr = len(branches)
    xs = [torch.randn(16, L, C, device=dev).bfloat16() for _ in branches]
    w1 = [torch.randn(C, C, k, device=dev) / (C * k) ** 0.5 for k, _ in branches]
    w2 = [torch.randn(C, C, k, device=dev) / (C * k) ** 0.5 for k, _ in branches]
    b = [torch.zeros(C, device=dev) for _ in branches]
    ds = [d for _, d in branches]
    ops.mrf_pairs_cl(xs, w1, b, w2, b, ds)
    trace.zero_()
    lib.vitsdec_debug_set_trace(trace.data_ptr())
    ops.mrf_pairs_cl(xs, w1, b, w2, b, ds)
    lib.vitsdec_debug_set_trace(None)
    torch.cuda.synchronize()
    t = trace.view(256, 12).cpu()
    n = int((t[:, 3] > 0).sum())
    t = t[:n].double()
    base = t[0, 0]
    print("branches %s  steps traced %d" % (branches, n))
    lo = (n // 2) // nbr * nbr
    for i in range(lo, min(n, lo + 2 * nbr)):
        print("  n %3d: c1 %7d-%7d  c2 %7d-%7d | epi1 %7d-%7d epi2 %7d-%7d | epi2 wait from %7d  epi1 wait from %7d  TMA issued %7d  c1 wait from %7d"
              % ((i,) + tuple(int(v - base) if v > 0 else -1 for v in t[i][:12])))
    last = torch.arange(nbr - 1, n, nbr)          # the step that carries the tile's epi2 stamps
    last = last[(last >= 4 * nbr) & (last < n - 2 * nbr)]
    per = (t[last[-1], 7] - t[last[0], 7]) / (len(last) - 1)
    s = slice(4 * nbr, n - 2 * nbr)
    print("  cycles/tile %.0f | per step: c1 issue %.0f  c2 issue %.0f  epi1 %.0f | epi2 %.0f | c1 end->epi1 start %.0f  epi1 end->c2 start %.0f"
          "  | last c2 end->epi2 start %.0f  c1 waits %.0f (acc1/TMA)  epi1 idle %.0f"
          % (per, (t[s, 1] - t[s, 0]).mean(), (t[s, 3] - t[s, 2]).mean(), (t[s, 5] - t[s, 4]).mean(),
             (t[last, 7] - t[last, 6]).mean(), (t[s, 4] - t[s, 1]).mean(), (t[s, 2] - t[s, 5]).mean(),
             (t[last, 6] - t[last, 3]).mean(), (t[s, 0] - t[s, 11]).mean(), (t[s, 4] - t[s, 9]).mean()))

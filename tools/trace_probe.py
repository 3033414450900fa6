"""Per-tile hand-off timeline of CTA 0 of conv_tc_kernel (clock64 stamps) for a few shapes."""
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
CASES = ((32, 220672, 1, 0), (32, 220672, 3, 0), (32, 220672, 3, 1), (64, 110336, 3, 0),
         (128, 55168, 3, 0), (128, 55168, 7, 0))
if len(sys.argv) > 1 and sys.argv[1] == "wide":
    CASES = ((128, 55168, 7, 0), (128, 55168, 11, 1), (256, 6896, 3, 0), (256, 6896, 11, 0))
for (C, L, k, use_res) in CASES:
    x = torch.randn(16, L, C, device=dev).bfloat16()
    r = torch.randn(16, L, C, device=dev).bfloat16() if use_res else None
    w = torch.randn(C, C, k, device=dev) / (C * k) ** 0.5
    b = torch.zeros(C, device=dev)
    ops.conv1d_cl(x, w, b, res=r, out_slope=0.1)  # warm
    trace.zero_()
    lib.vitsdec_debug_set_trace(trace.data_ptr())
    ops.conv1d_cl(x, w, b, res=r, out_slope=0.1)
    lib.vitsdec_debug_set_trace(None)
    torch.cuda.synchronize()
    t = trace.view(256, 12).cpu()
    n = int((t[:, 3] > 0).sum())
    t = t[:n]
    base = int(t[0, 0])
    print("C=%d k=%d res=%d tiles traced %d" % (C, k, use_res, n))
    print("  cols: prod_acq mma_accfree mma_afull mma_issued epi_accfull epi_tmemld epi_accum epi_stored (cycles since first)")
    for i in list(range(max(6, n // 2), min(n, n // 2 + 3))):
        print("  tile %3d: " % i + " ".join("%8d" % (int(v) - base) for v in t[i][:10]))
    if n > 20:
        d = (t[n - 5, 7] - t[10, 7]).item() / (n - 15)
        print("  steady-state cycles per tile: %.0f;  per-tile means: a_full wait %.0f  mma issue %.0f  epi wait->ld %.0f  "
              "accum %.0f  release %.0f  next-coords+loads %.0f  store %.0f" % (d, (t[10:n-5, 2] - t[10:n-5, 1]).float().mean(), (t[10:n-5, 3] - t[10:n-5, 2]).float().mean(),
                                          (t[10:n-5, 5] - t[10:n-5, 4]).float().mean(), (t[10:n-5, 6] - t[10:n-5, 5]).float().mean(),
                                          (t[10:n-5, 8] - t[10:n-5, 6]).float().mean(), (t[10:n-5, 9] - t[10:n-5, 8]).float().mean(),
                                          (t[10:n-5, 7] - t[10:n-5, 9]).float().mean()))

"""Time the native flow alone (reverse pass) at a few sizes; with --once runs a single pass (for ncu launch lists)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitsdec
once = "--once" in sys.argv
F = vitsdec.ResidualCouplingBlock(192, 192, 5, 1, 4, gin_channels=256)
with torch.no_grad():
    for n, p in F.named_parameters():
        if n.endswith("post.weight"):
            p.uniform_(-0.07, 0.07)
F = F.to("cuda:0").eval()
F.assume_frozen = True
for pdl in ((1,) if once else (0, 1)):
  F.set_option("pdl", pdl)
  for B, T in (((16, 862),) if once else ((1, 173), (1, 862), (16, 862))):
      x = torch.randn(B, 192, T, device="cuda:0"); g = torch.randn(B, 256, 1, device="cuda:0"); m = torch.ones(B, 1, T, device="cuda:0")
      with torch.no_grad():
          for _ in range(1 if once else 5): F(x, m, g=g, reverse=True)
          if once:
              torch.cuda.synchronize(); break
          e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
          torch.cuda.synchronize(); e0.record()
          for _ in range(30): F(x, m, g=g, reverse=True)
          e1.record(); torch.cuda.synchronize()
      print("flow reverse pdl=%d B=%d T=%d: %.3f ms" % (pdl, B, T, e0.elapsed_time(e1) / 30))

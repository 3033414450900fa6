"""Decode latency of small batches (BASELINE config 2: batch 1, 2 s) under option variants, one process, one box."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitsdec
cargs, ckw = vitsdec.generator_args()
torch.manual_seed(1)
G = vitsdec.Generator(*cargs, **ckw).to("cuda:0").eval()
G.assume_frozen = True
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for (B, T) in ((1, 173), (1, 862), (4, 173), (16, 862)):
    z = torch.randn(B, cargs[0], T, device="cuda:0")
    g = torch.randn(B, ckw["gin_channels"], 1, device="cuda:0")
    out = []
    for rnd in range(2):
        for pdl in (0, 1, 2):
            G.set_option("pdl", pdl)
            with torch.no_grad():
                for _ in range(6):
                    y = G(z, g)
                torch.cuda.synchronize()
                n = 50 if B * T < 5000 else 10
                e0.record()
                for _ in range(n):
                    y = G(z, g)
                e1.record()
                torch.cuda.synchronize()
            out.append("pdl=%d %.4f ms" % (pdl, e0.elapsed_time(e1) / n))
    print("B=%d T=%d: " % (B, T) + " | ".join(out))
    if rnd == 1 and B == 1 and T == 173:
        ref = y.clone()
G.set_option("pdl", 0)
z = torch.randn(2, cargs[0], 100, device="cuda:0"); g = torch.randn(2, ckw["gin_channels"], 1, device="cuda:0")
with torch.no_grad():
    a = G(z, g).clone(); G.set_option("pdl", 1); b = [G(z, g).clone() for _ in range(4)][-1]; G.set_option("pdl", 2); c = [G(z, g).clone() for _ in range(4)][-1]
print("bitwise equal across pdl modes:", bool(torch.equal(a, b) and torch.equal(a, c)))

"""Decode time with serial vs concurrent MRF branches over batch sizes (chooses the auto threshold of option par)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitsdec
cargs, ckw = vitsdec.generator_args()
torch.manual_seed(0)
G = vitsdec.Generator(*cargs, **ckw).to("cuda:0").eval()
G.assume_frozen = True
for B, T in ((1, 173), (1, 431), (1, 862), (2, 862), (4, 862), (8, 862), (16, 862)):
    z = torch.randn(B, cargs[0], T, device="cuda:0"); g = torch.randn(B, 256, 1, device="cuda:0")
    res = []
    for par in (0, 2):
        G.set_option("par", par)
        with torch.no_grad():
            for _ in range(5): G(z, g)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            n = 30
            for _ in range(n): G(z, g)
            e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / n)
    print("B=%d T=%d frames=%d  serial %.3f ms  concurrent %.3f ms  ratio %.3f" % (B, T, B * T, res[0], res[1], res[1] / res[0]))

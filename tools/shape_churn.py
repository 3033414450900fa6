"""Latency of decodes whose shape changes every call (TTS serving) vs repeated shapes: plan building, graph capture."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitsdec
cargs, ckw = vitsdec.generator_args()
torch.manual_seed(0)
G = vitsdec.Generator(*cargs, **ckw).to("cuda:0").eval()
G.assume_frozen = True
g = torch.randn(1, 256, 1, device="cuda:0")
with torch.no_grad():
    G(torch.randn(1, 192, 50, device="cuda:0"), g); torch.cuda.synchronize()
    for graph in (1, 0):
        G.set_option("graph", graph)
        first, again = [], []
        for T in range(100 + 200 * graph, 400 + 200 * graph, 13):   # fresh shapes for each mode
            z = torch.randn(1, 192, T, device="cuda:0")
            torch.cuda.synchronize(); t0 = time.perf_counter(); G(z, g); torch.cuda.synchronize(); first.append(time.perf_counter() - t0)
            t0 = time.perf_counter(); G(z, g); torch.cuda.synchronize(); again.append(time.perf_counter() - t0)
        print("graph=%d  new shape: %.2f ms   same shape again: %.2f ms (wall clock per decode, batch 1, 100-400 frames)"
              % (graph, 1e3 * sum(first) / len(first), 1e3 * sum(again) / len(again)))

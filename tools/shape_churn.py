"""Serving with changing shapes, DEFAULT module settings (no assume_frozen): wall-clock latency per batch-1 decode for
(a) a shape seen for the first time (plan build + plain launches), (b) the same shape again, and (c) a mixed-shape loop
(requests drawn from a pool of lengths: plans are cached, a plan replays as a CUDA graph from its third use on)."""
import os, sys, time, random, statistics, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitsdec
cargs, ckw = vitsdec.generator_args()
torch.manual_seed(0)
G = vitsdec.Generator(*cargs, **ckw).to("cuda:0").eval()
g = torch.randn(1, 256, 1, device="cuda:0")


def timed(z):
    torch.cuda.synchronize(); t0 = time.perf_counter(); G(z, g); torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0)


with torch.no_grad():
    G(torch.randn(1, 192, 50, device="cuda:0"), g); torch.cuda.synchronize()
    for graph in (1, 0):
        G.set_option("graph", graph)
        first, again = [], []
        for T in range(100 + 200 * graph, 400 + 200 * graph, 13):   # fresh shapes for each mode
            z = torch.randn(1, 192, T, device="cuda:0")
            first.append(timed(z)); again.append(timed(z))
        print("graph=%d  new shape: %.2f ms   same shape again: %.2f ms (wall clock per decode, batch 1, 100-400 frames)"
              % (graph, sum(first) / len(first), sum(again) / len(again)))
    G.set_option("graph", 1)
    random.seed(1)
    pool = [random.randrange(80, 900) for _ in range(24)]       # 1-10 s utterances
    zs = {T: torch.randn(1, 192, T, device="cuda:0") for T in pool}
    lat = [(T, timed(zs[T])) for T in (random.choice(pool) for _ in range(600))]
    cold, warm = [l for _, l in lat[:100]], [l for _, l in lat[300:]]
    print("mixed-shape loop, 24 lengths in 80-900 frames, batch 1: first 100 requests mean %.2f ms; requests 300-600: mean %.3f "
          "median %.3f p95 %.3f ms (wall clock incl. Python and the synchronize)"
          % (statistics.mean(cold), statistics.mean(warm), statistics.median(warm), sorted(warm)[int(0.95 * len(warm))]))
    print("graph_failed plans:", G.get_option("graph_failed"))

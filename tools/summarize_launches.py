"""Per-launch table of one decode step from an ncu CSV with gpu__time_duration / dram bytes / tensor-pipe metrics.
Writes the text summary committed under profiles/ and (optionally) the per-step traffic JSON bench.py reads."""
import csv
import json
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    by = {}
    order = []
    for r in csv.DictReader(lines):
        k = r["ID"]
        if k not in by:
            by[k] = {"name": r["Kernel Name"].split("(")[0].replace("void ", "")}
            order.append(k)
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            v = float("nan")
        u = r["Metric Unit"]
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        elif m.startswith("dram__bytes"):
            v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}[u]
        by[k][m] = v
    return [by[k] for k in order]


def main(path, out_json=None):
    rows = load(path)
    starts = [i for i, r in enumerate(rows) if "pack_z" in r["name"]] + [len(rows)]
    n_step = starts[1] - starts[0]                       # launches of one decode
    full = [i for i in range(len(starts) - 1) if starts[i + 1] - starts[i] == n_step]
    start = starts[full[-1]]                             # the last complete decode of the capture
    step = rows[start:start + n_step]
    T = "gpu__time_duration.sum"
    tot = sum(r[T] for r in step)
    print("# one decode step (16 x 10 s): %d launches, %.1f us summed device time (ncu, cold-cache, serialised)" % (len(step), tot))
    print("# idx kernel                                   us    share   dram_rd_MB dram_wr_MB  tensor_pipe_%")
    conv = {"us": 0.0, "rd": 0.0, "wr": 0.0, "n": 0}
    for i, r in enumerate(step):
        rd, wr = r.get("dram__bytes_read.sum", 0), r.get("dram__bytes_write.sum", 0)
        tp = float("nan")   # tensor-pipe active %: the first of the requested spellings this ncu build answers
        for key in r:
            if "pipe_tensor" in key and r[key] == r[key]:
                tp = r[key]
                break
        print("%3d %-38s %8.1f %6.2f%% %10.1f %10.1f %10.1f" % (i, r["name"][:38], r[T], 100 * r[T] / tot, rd, wr, tp))
        if "conv_tc" in r["name"] or "conv_pair" in r["name"] or "conv_mrf" in r["name"]:   # conv_pair matches conv_pairf too, conv_mrf both conv_mrfp and conv_mrf128
            conv["us"] += r[T]; conv["rd"] += rd; conv["wr"] += wr; conv["n"] += 1
    print("# tcgen05 conv kernels: %d launches, %.1f us (%.1f%% of the step), DRAM read %.0f MB + write %.0f MB per step"
          % (conv["n"], conv["us"], 100 * conv["us"] / tot, conv["rd"], conv["wr"]))
    if out_json:
        import importlib
        import os
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        digest = importlib.import_module("personalized_text-to-speech_b200.build")._digest()   # the kernels captured
        json.dump({"conv_launches_per_step": conv["n"], "conv_dram_bytes_per_step": (conv["rd"] + conv["wr"]) * 1e6,
                   "csrc_digest": digest,
                   "conv_share_of_step": conv["us"] / tot, "source": path.split("/")[-1],
                   "workload": "16 x 10 s decode"}, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)

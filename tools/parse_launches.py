"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of one decode step per conv layer."""
import csv
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    out = []
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else (v * 1e6 if u == "s" else v))
        out.append((row["Kernel Name"].split("(")[0], v))
    return out


def main(path):
    rows = load(path)
    start = next(i for i, (n, _) in enumerate(rows) if "pack_z" in n)
    nxt = next((i for i, (n, _) in enumerate(rows) if "pack_z" in n and i > start), len(rows))
    per_step = nxt - start if nxt - start in (72, 80) else 80
    step = rows[start:start + per_step]
    names = ["pack_z", "cond", "conv_pre"]
    ks = [3, 7, 11]
    for st in range(4):
        names.append("ups%d" % st)
        for k in ks:
            for m in range(3):
                names.append("s%d k%-2d c1.%d" % (st, k, m))
                if per_step == 80 or m < 2:
                    names.append("s%d k%-2d c2.%d" % (st, k, m))
        if per_step == 72:
            names.append("mrf%d" % st)
    names.append("conv_post")
    tot = sum(v for _, v in step)
    print("step total %.1f us over %d launches" % (tot, len(step)))
    stage = {}
    for (kn, v), nm in zip(step, names):
        key = nm.split()[0] if nm.startswith("s") else nm
        stage[key] = stage.get(key, 0) + v
    for st in range(4):
        line = "stage %d: " % st
        for k in ks:
            vals = [v for (kn, v), nm in zip(step, names) if nm.startswith("s%d k%-2d" % (st, k))]
            line += " k%-2d [%s] = %.0f |" % (k, " ".join("%.0f" % v for v in vals), sum(vals))
        print(line + "  total %.0f us (ups %.0f, fused mrf %.0f)" % (stage["s%d" % st] + stage.get("mrf%d" % st, 0),
                                                                    stage["ups%d" % st], stage.get("mrf%d" % st, 0)))
    print("conv_pre %.1f  conv_post %.1f  pack_z %.1f cond %.1f" % (stage["conv_pre"], stage["conv_post"],
                                                                    stage["pack_z"], stage["cond"]))


if __name__ == "__main__":
    main(sys.argv[1])

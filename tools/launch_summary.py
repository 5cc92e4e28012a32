"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per kernel and per stage."""
import collections
import csv
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    out = []
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        out.append((row["Kernel Name"].split("(")[0].replace("void ", "").replace("unnamed>::", ""), v, row["Grid Size"]))
    return out


def main():
    rows = load(sys.argv[1])
    agg = collections.OrderedDict()
    for n, v, g in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:60]:60s} n={a[0]:4d} total={a[1] / 1e3:9.3f} ms share={a[1] / tot * 100:5.1f}%")
    print(f"total {tot / 1e3:.3f} ms over {len(rows)} launches")
    # per-stage: split at convtr launches (the conv launch directly before an act run of a new size)
    if "-v" in sys.argv:
        for i, (n, v, g) in enumerate(rows):
            if "conv_umma" in n or "act1d" in n:
                print(i, n[:30], g, f"{v:9.1f} us")
    # conv / act per distinct grid (one grid per stage)
    per = collections.OrderedDict()
    for n, v, g in rows:
        if "act1d_c8t" in n:
            p = per.setdefault(("act", g), [0, 0.0]); p[0] += 1; p[1] += v
    for k, p in per.items():
        print(f"act grid {k[1]:>16s}: n={p[0]:3d} avg={p[1] / p[0]:8.1f} us total={p[1] / 1e3:7.2f} ms")


if __name__ == "__main__":
    main()

"""Summarise an `ncu --csv --metrics gpu__time_duration.sum` launch list: the LAST n launches grouped by kernel name."""
import csv, sys, re
from collections import OrderedDict
path, n = sys.argv[1], int(sys.argv[2])
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
r = list(csv.DictReader(lines))
r = [x for x in r if x.get("Metric Name") == "gpu__time_duration.sum"]
r = r[-n:]
agg = OrderedDict()
for x in r:
    k = re.sub(r"\(.*", "", x["Kernel Name"])
    v = float(x["Metric Value"].replace(",", ""))
    u = x["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else v if u in ("us", "usecond") else v * 1e3
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += v
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:60s} n={c:4d} total={t:9.1f} us  avg={t / c:7.1f} us")
print(f"total {sum(t for _, t in agg.values()):.1f} us over {sum(c for c, _ in agg.values())} launches")

"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import gzip
import re
import sys

path = sys.argv[1]
op = gzip.open if path.endswith(".gz") else open
with op(path, "rt") as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    short = re.sub(r"\(.*", "", row["Kernel Name"])[:72]
    agg[short][0] += 1
    agg[short][1] += v
tot = sum(v for _, v in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print("%10.3f ms %5.1f%% n=%5d avg=%9.1f us  %s" % (t / 1e3, 100 * t / tot, n, t / n, k))
print("total %.3f ms over %d launches" % (tot / 1e3, sum(n for n, _ in agg.values())))

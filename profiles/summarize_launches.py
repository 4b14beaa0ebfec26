"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python profiles/summarize_launches.py launches.csv[.gz] > summary.md"""
import collections
import csv
import gzip
import re
import sys

path = sys.argv[1]
opener = gzip.open if path.endswith(".gz") else open
with opener(path, "rt") as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
total = 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = re.sub(r"<.*", "", name)[:80]
    agg[name][0] += 1
    agg[name][1] += v
    total += v
print("| launches | total us | share | kernel |\n|---:|---:|---:|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| %d | %.1f | %.1f%% | `%s` |" % (n, t, 100 * t / total, k))
own = sum(t for k, (n, t) in agg.items() if re.search(r"flowk::|tc::|attn_train::", k))
print("\nflowk kernels: %.1f%% of the device time (the rest: ATen / cuBLAS / MAGMA glue)" % (100 * own / total))
print("\ntotal: %d launches, %.1f us (cold-cache, serialised under ncu: compare shares, not absolutes)"
      % (sum(a[0] for a in agg.values()), total))

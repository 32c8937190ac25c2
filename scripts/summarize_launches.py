"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel.
usage: python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.md"""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
total = 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
    agg[name][0] += 1
    agg[name][1] += v
    total += v
print("# ncu launch list summary: %s" % path)
print("\nPer-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's own CUDA-event")
print("numbers, not absolutes.  Total device time of the captured launches: %.3f ms\n" % total)
print("| kernel | launches | total ms | share |")
print("|---|---:|---:|---:|")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.3f | %.3f |" % (k[:90], n, ms, ms / total))

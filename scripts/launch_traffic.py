"""From an ncu launch list with gpu__time_duration.sum, dram__bytes_read.sum and dram__bytes_write.sum (one bench
step between two adam_kernel launches): per-kernel table (markdown, stdout) and the per-class DRAM traffic per
launch that bench.py reports as roofline.traffic (JSON, second argument).
usage: python scripts/launch_traffic.py gpurun_out/launches.csv profiles/r01_traffic.json > profiles/<name>.md"""
import collections
import csv
import json
import re
import sys

path, out_json = sys.argv[1], sys.argv[2]
lines = [l for l in open(path) if not l.startswith("==")]
by_id = collections.OrderedDict()
for r in csv.DictReader(lines):
    d = by_id.setdefault(r["ID"], {"name": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("dlio::", ""),
                                   "grid": r["Grid Size"]})
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        d["ms"] = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[u]
    elif u in ("byte", "Kbyte", "Mbyte", "Gbyte"):
        d[r["Metric Name"]] = v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    else:
        d[r["Metric Name"]] = v            # instruction counts, percentages
items = list(by_id.values())
adam = [i for i, d in enumerate(items) if "adam_kernel" in d["name"]]
# one optimizer step = a run of consecutive adam launches (one per parameter group)
ends = [i for k, i in enumerate(adam) if k + 1 == len(adam) or adam[k + 1] != i + 1]
step = items[ends[0] + 1:ends[1] + 1]
total = sum(d["ms"] for d in step)
agg = collections.OrderedDict()
for d in step:
    a = agg.setdefault(d["name"][:70], [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += d["ms"]
    a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    a[3] += d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0) * d["ms"]
print("# ncu launch list of one train step: %s\n" % path)
print("Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's CUDA-event numbers,")
print("not absolutes.  %d launches, %.3f ms of device time, DRAM traffic %.2f GB.\n" % (
    len(step), total, sum(a[2] for a in agg.values()) / 1e9))
print("| kernel | launches | total ms | share | DRAM GB | GB/s | issue active % (time-weighted) |")
print("|---|---:|---:|---:|---:|---:|---:|")
for k, (n, ms, by, ia) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.3f | %.3f | %.3f | %.0f | %.0f |" % (k, n, ms, ms / total, by / 1e9,
                                                          by / 1e9 / (ms * 1e-3) if ms else 0, ia / ms if ms else 0))
# classes as bench.py names them: conv_tc_kernel launches before the step's first wgrad are forward, later ones dgrad
first_wgrad = next(i for i, d in enumerate(step) if re.search(r"wgrad_tc2?_kernel", d["name"]))
cls = {"conv_fwd_tc": [], "conv_dgrad_tc": [], "conv_wgrad_tc": []}
for i, d in enumerate(step):
    if re.search(r"wgrad_tc2?_kernel", d["name"]):
        cls["conv_wgrad_tc"].append(d)
    elif re.search(r"conv_tc2?_kernel", d["name"]):
        cls["conv_fwd_tc" if i < first_wgrad else "conv_dgrad_tc"].append(d)
res = {}
for k, ds in cls.items():
    by = sum(d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0) for d in ds)
    res[k] = {"launches_per_step": len(ds), "dram_bytes_per_launch": by / max(1, len(ds)), "ms_per_step_under_ncu": sum(d["ms"] for d in ds),
              "source": path}
json.dump(res, open(out_json, "w"), indent=1)
print("\nPer class (bench.py roofline.traffic = DRAM bytes per launch): " + json.dumps(
    {k: {"launches": v["launches_per_step"], "MB_per_launch": round(v["dram_bytes_per_launch"] / 1e6, 1)} for k, v in res.items()}))

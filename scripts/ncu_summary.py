"""Markdown table of the headline metrics of every launch in an .ncu-rep (needs `ncu` on PATH).
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.md"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
cols = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "time"),
        ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
        ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe % (elapsed)"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn smem"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %")]
idx = [(hdr.index(c), n) for c, n in cols if c in hdr]
print("# ncu --set full summary: %s\n" % rep)
print("| " + " | ".join("%s%s" % (n, (" [%s]" % units[i]) if units[i] else "") for i, n in idx) + " |")
print("|" + "---|" * len(idx))
for r in data:
    cells = []
    for i, n in idx:
        v = r[i]
        if n == "kernel":
            v = v.split("(")[0]
        else:
            try:
                v = "%.3f" % float(v.replace(",", ""))
            except ValueError:
                pass
        cells.append(v)
    print("| " + " | ".join(cells) + " |")

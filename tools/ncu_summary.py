"""Summarise an .ncu-rep (raw page) per kernel launch: duration, DRAM bytes, issue utilisation, top stalls."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
def col(name): return hdr.index(name) if name in hdr else -1
base = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in data:
    print("==", r[col("Kernel Name")][:90], "grid", r[col("Grid Size")] if col("Grid Size") >= 0 else "")
    for b in base:
        i = col(b)
        if i >= 0: print(f"   {b:72s} {r[i]:>16s} {units[i]}")
    st = sorted(((float(r[hdr.index(s)].replace(',', '')), s) for s in stalls), reverse=True)[:5]
    print("   top stalls:", ", ".join(f"{s.split('stalled_')[1].split('_per_issue')[0]}={v:.1f}" for v, s in st))

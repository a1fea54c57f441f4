"""Per-source-line instruction / stall-sample totals from an .ncu-rep (needs -lineinfo + --import-source).
usage: ncu_lines.py report.ncu-rep kernel_regex [top] [sort: inst|samples] [launch_index]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
key = sys.argv[4] if len(sys.argv) > 4 else "inst"
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern]
if len(sys.argv) > 5:
    cmd += ["--launch-skip", sys.argv[5], "--launch-count", "1"]
raw = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
agg, cur_file, hdr = {}, None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; ie = hdr.index("Instructions Executed"); ns = hdr.index("# Samples"); continue
    if hdr is None or len(r) <= ie:
        continue
    if r[2] == "-":
        a = agg.setdefault((cur_file, int(r[0]), r[1].strip()), [0, 0])
        a[0] += int(r[ie] or 0); a[1] += int(r[ns] or 0)
tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
print(f"total warp-instructions {tot_i}, samples {tot_s}")
k = 0 if key == "inst" else 1
for (f, ln, src), (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][k])[:top]:
    print(f"inst {i:>11d} {100 * i / max(tot_i, 1):5.1f}%  smp {100 * s / max(tot_s, 1):5.1f}%  {f}:{ln:<4d} {src[:95]}")

"""Per-source-line instruction / stall-sample totals from an .ncu-rep (needs -lineinfo + --import-source)."""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
agg = {}
cur_file = None
hdr = None
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
    if r[2] == "-":   # a source line (aggregated over its SASS)
        key = (cur_file, int(r[0]), r[1].strip())
        a = agg.setdefault(key, [0, 0])
        a[0] += int(r[ie] or 0); a[1] += int(r[ns] or 0)
tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
print(f"total warp-instructions {tot_i}, samples {tot_s}")
for (f, ln, src), (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{i:>12d} {100 * i / max(tot_i, 1):5.1f}%  smp {100 * s / max(tot_s, 1):5.1f}%  {f}:{ln:<4d} {src[:100]}")

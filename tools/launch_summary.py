"""Per-kernel totals and shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr, data = rows[hi], rows[hi + 1:]
kn, mv, gs, bs = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= mv:
        continue
    a = agg.setdefault(r[kn] + " " + r[gs], [0, 0.0, r[gs], r[bs]])      # launches of different grids are different rows
    a[0] += 1; a[1] += float(r[mv].replace(",", "")) / 1e3
own = {k: v for k, v in agg.items() if "at::native" not in k and "elementwise" not in k and "nccl" not in k.lower()}
tot = sum(v[1] for v in own.values())
print(f"{'launches':>8s} {'avg us':>9s} {'share':>7s}  kernel (grid x block)")
for k, (n, t, g, b) in own.items():
    print(f"{n:8d} {t / n:9.2f} {100 * t / tot:6.1f}%  {k.replace('b200det::', '')[:90]} {g}x{b}")
k1 = next((k for k in own if "yolo_decode" in k), None)
if k1:
    print(f"sum of own kernels per K1 launch (= per step): {tot / own[k1][0]:.1f} us")

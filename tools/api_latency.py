"""Wall-clock latency of the public API call (host overhead + GPU + the one count sync) for small batches."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import objectdetectionpl_b200 as od
from objectdetectionpl_b200 import synth
dev = torch.device("cuda:0")
CASES = ((1, 20, "uniform", 0.5), (1, 80, "sparse", 0.25), (8, 80, "sparse", 0.25), (64, 80, "uniform", 0.5))
if os.environ.get("API_LATENCY_ONLY"):
    CASES = tuple(c for c in CASES if str(c[0]) in os.environ["API_LATENCY_ONLY"].split(","))
for B, C, mode, thr in CASES:
    lv = [t.to(dev) for t in synth.yolo_planar(B, 3, C, [80, 40, 20], 640, 3, conf_mode=mode, v5_view=True, tie_free=False)]
    kw = dict(compat=(mode == "uniform"), conf_thres=thr)
    for _ in range(5):
        out = od.non_max_suppression(None, lv, **kw)     # keep the result like the timed loop does (one more live output block)
    torch.cuda.synchronize()
    n = 50
    per = []
    t0 = time.perf_counter()
    for _ in range(n):
        t1 = time.perf_counter()
        out = od.non_max_suppression(None, lv, **kw)
        per.append((time.perf_counter() - t1) * 1e6)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    t0 = time.perf_counter()
    for _ in range(n):
        od.yolo_nms_raw(lv, 3, od.YOLO_FORCED_CONF_THRES if mode == "uniform" else thr)
    torch.cuda.synchronize()
    dr = (time.perf_counter() - t0) / n
    print(f"B={B} C={C} {mode}: non_max_suppression {dt * 1e6:.0f} us/call, yolo_nms_raw (no sync, no list) {dr * 1e6:.0f} us/call, "
          f"kept {sum(o.shape[0] for o in out if o is not None)}; per-call median {sorted(per)[n // 2]:.0f} us, max {max(per):.0f} us, "
          f"calls above 2x the median: {sum(p > 2 * sorted(per)[n // 2] for p in per)}")

"""Where the public-API call spends its host time (B=64 headline shapes): phases of _yolo_nms_planned with perf_counter,
for results that are dropped at once and for results the caller keeps (two output blocks alternate in the allocator)."""
import os, sys, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import objectdetectionpl_b200 as od
from objectdetectionpl_b200 import synth, postprocess as P, _lib as L

dev = torch.device("cuda:0")
lv = [t.to(dev) for t in synth.yolo_planar(64, 3, 80, [80, 40, 20], 640, 3, v5_view=True, tie_free=False)]
for _ in range(5):
    od.non_max_suppression(None, lv)
torch.cuda.synchronize()


def wall(fn, n=100):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


keep = [None]
def keepit():
    keep[0] = od.non_max_suppression(None, lv)
print("drop result   : %.0f us/call" % wall(lambda: od.non_max_suppression(None, lv)))
print("keep result   : %.0f us/call" % wall(keepit))
print("raw (no sync) : %.0f us/call" % wall(lambda: od.yolo_nms_raw(lv, 3)))

# phase breakdown: perf_counter stamps postprocess._yolo_nms_planned takes at its phase boundaries when P._trace is a list
for pattern, fn in (("drop", lambda: od.non_max_suppression(None, lv)), ("keep", keepit)):
    P._trace = tr = []
    torch.cuda.synchronize()
    for _ in range(200):
        fn()
    torch.cuda.synchronize()
    P._trace = None
    calls = [tr[i:i + 6] for i in range(0, len(tr), 6)][10:]
    m = [statistics.median(c[i + 1] - c[i] for c in calls) * 1e6 for i in range(5)]
    gap = statistics.median(calls[i + 1][0] - calls[i][5] for i in range(len(calls) - 1)) * 1e6
    print(pattern, "prologue %.1f | launch (one ctypes call) %.1f | views + next buffers %.1f | wait for the counts event %.1f | "
          "finish_views %.1f | between calls (return, drop, plan lookup, entry) %.1f  (us, medians)" % (tuple(m) + (gap,)))
ev0 = torch.cuda.Event()
torch.cuda.synchronize()
lat = []
for _ in range(200):
    ev0.record(); t0 = time.perf_counter(); ev0.synchronize(); lat.append(time.perf_counter() - t0)
print("event synchronize right after its record on an idle stream: %.1f us median" % (statistics.median(lat) * 1e6))

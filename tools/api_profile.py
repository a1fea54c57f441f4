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

# phase breakdown (drop pattern)
orig = P._yolo_nms_planned
marks = []
def timed(plan, predictions, d, B, n_pad, dev_, return_index):
    t0 = time.perf_counter()
    for i, t in enumerate(predictions):
        d.head[i] = t.data_ptr()
    with torch.cuda.device(dev_):
        ws = L.workspace(plan.ws_bytes, dev_)
        rows = torch.empty((B * n_pad, 7), dtype=torch.float32, device=dev_)
        meta = torch.empty((2 * B + 1,), dtype=torch.int32, device=dev_)
        mp = meta.data_ptr()
        t1 = time.perf_counter()
        L.check(plan.fn(plan.dref, ws.data_ptr(), ws.numel(), rows.data_ptr(), None, mp, mp + 4 * B, plan.host.data_ptr(),
                        plan.event.cuda_event, L.stream_ptr(dev_)), "x")
        t2 = time.perf_counter()
        plan.event.synchronize()
        t3 = time.perf_counter()
    meta_h = plan.host.tolist()
    counts, total = meta_h[:B], meta_h[2 * B]
    t4 = time.perf_counter()
    parts = torch.ops.aten.unsafe_split_with_sizes.default(rows[:total], counts)
    t5 = time.perf_counter()
    out = [p if k else None for p, k in zip(parts, counts)]
    t6 = time.perf_counter()
    marks.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5))
    return out
P._yolo_nms_planned = timed
for pattern, fn in (("drop", lambda: od.non_max_suppression(None, lv)), ("keep", keepit)):
    marks.clear()
    torch.cuda.synchronize()
    for _ in range(100):
        fn()
    torch.cuda.synchronize()
    m = [statistics.median(x[i] for x in marks[10:]) * 1e6 for i in range(6)]
    print(pattern, "alloc %.1f | launch %.1f | wait %.1f | tolist %.1f | split %.1f | list %.1f  (us, medians)" % tuple(m))

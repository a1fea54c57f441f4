"""CUDA-event time of the sort / NMS / emit stages for the kernel variants selectable by environment variables
(B200DET_NMS=half2|tab, B200DET_SORT=...), on the headline, YOLOv3-416 and dense-crowd workloads.
    python tools/nms_ab_timing.py [--iters 50] [--modes half2,tab]"""
import argparse, ctypes, json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectdetectionpl_b200 import _lib as L, synth
from objectdetectionpl_b200.postprocess import _yolo_desc

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=50)
ap.add_argument("--modes", default="half2,tab")
ap.add_argument("--env", default="B200DET_NMS")
ap.add_argument("--workloads", default="headline,cfg2,crowd")
a = ap.parse_args()
dev = torch.device("cuda:0")
lib = L.load()
WL = {
    "headline": lambda: (synth.yolo_planar(64, 3, 80, [80, 40, 20], 640, 1234, v5_view=True, tie_free=False), -0.0151),
    "cfg2": lambda: (synth.yolo_planar(64, 3, 80, [13, 26, 52], 416, 2, tie_free=False), -0.0151),
    "crowd": lambda: (synth.yolo_crowd(64, 3, 5, [160, 80, 40], 1280, seed=5), 0.001),
    "cfg1": lambda: (synth.yolo_planar(1, 3, 20, [80, 40, 20], 640, 1, v5_view=True, tie_free=False), -0.0151),
}
for name in a.workloads.split(","):
    lv_cpu, thr = WL[name]()
    lv = [t.to(dev) for t in lv_cpu]
    d = _yolo_desc(lv, 3, thr, 0.4, None, None, None)
    n, n_pad = ctypes.c_int32(), ctypes.c_int32()
    L.check(lib.b200det_yolo_num_candidates(ctypes.byref(d), ctypes.byref(n), ctypes.byref(n_pad)))
    wsb = lib.b200det_yolo_workspace_bytes(ctypes.byref(d))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    B = lv[0].shape[0]
    rows = torch.empty((B, n_pad.value, 7), device=dev)
    cnt = torch.empty((B,), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    dref, wp = ctypes.byref(d), ws.data_ptr()
    ref_rows = None
    for mode in a.modes.split(","):
        os.environ[a.env] = mode
        def step(ev=None):
            L.check(lib.b200det_yolo_stage_reset(dref, wp, wsb, st))
            if ev: ev[0].record()
            L.check(lib.b200det_yolo_stage_decode(dref, wp, wsb, st))
            if ev: ev[1].record()
            L.check(lib.b200det_yolo_stage_sort(dref, wp, wsb, st))
            if ev: ev[2].record()
            L.check(lib.b200det_yolo_stage_nms(dref, wp, wsb, st))
            if ev: ev[3].record()
            L.check(lib.b200det_yolo_stage_emit(dref, wp, wsb, rows.data_ptr(), None, cnt.data_ptr(), st))
            if ev: ev[4].record()
        for _ in range(5):
            step()
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(a.iters)]
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for e in evs:
            step(e)
        torch.cuda.synchronize()
        t0.record()
        for _ in range(a.iters):
            step()
        t1.record()
        torch.cuda.synchronize()
        med = lambda k: statistics.median(e[k].elapsed_time(e[k + 1]) for e in evs) * 1e3
        cur = (rows.clone(), cnt.clone())
        same = None
        if ref_rows is not None:
            same = bool(torch.equal(cur[1], ref_rows[1])) and all(
                torch.equal(cur[0][b, :k], ref_rows[0][b, :k]) for b, k in enumerate(cur[1].tolist()))
        else:
            ref_rows = cur
        print(json.dumps({"workload": name, a.env: mode, "decode_us": round(med(0), 1), "sort_us": round(med(1), 1),
                          "nms_us": round(med(2), 1), "emit_us": round(med(3), 1),
                          "step_us_no_events": round(t0.elapsed_time(t1) / a.iters * 1e3, 1), "kept": int(cnt.sum()),
                          "identical_to_first_mode": same}))
    os.environ.pop(a.env, None)
    del lv, ws, rows

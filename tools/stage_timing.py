"""Per-stage CUDA-event timing of the YOLO pipeline (development aid; bench.py is the contract)."""
import argparse
import ctypes
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from objectdetectionpl_b200 import _lib as L, synth
from objectdetectionpl_b200.postprocess import _yolo_desc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--classes", type=int, default=80)
    ap.add_argument("--img", type=int, default=640)
    ap.add_argument("--model", default="yolov5")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--conf-mode", default="uniform")
    ap.add_argument("--conf-thres", type=float, default=-0.0151)
    ap.add_argument("--layout", default="planar", choices=["planar", "channels_last"])
    ap.add_argument("--crowd", action="store_true", help="BASELINE config 5 shard: synth.yolo_crowd 1280x1280, 5 classes, conf_thres 0.001")
    ap.add_argument("--trace-sort", action="store_true", help="per-phase %globaltimer trace of the cluster sort")
    ap.add_argument("--trace-nms", action="store_true", help="per-phase %globaltimer trace of the NMS kernel")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    if a.crowd:
        a.img, a.classes, a.conf_thres = 1280, 5, 0.001
    grids = synth.grids_for(a.model, a.img)
    g = torch.Generator(device=dev).manual_seed(1)
    levels = []
    if a.crowd:
        levels = [t.to(dev) for t in synth.yolo_crowd(a.batch, 3, 5, grids, 1280, seed=5)]
        grids = []
    for G in grids:   # generated on device for speed: same distributions as synth.yolo_planar
        t = torch.empty(a.batch, 3, 5 + a.classes, G, G, device=dev)
        t[:, :, 0:2] = torch.rand(a.batch, 3, 2, G, G, device=dev, generator=g) * a.img
        t[:, :, 2:4] = 8 + torch.rand(a.batch, 3, 2, G, G, device=dev, generator=g) * (a.img / 4 - 8)
        if a.conf_mode == "uniform":
            t[:, :, 4] = torch.rand(a.batch, 3, G, G, device=dev, generator=g)
        else:
            t[:, :, 4] = torch.sigmoid(torch.randn(a.batch, 3, G, G, device=dev, generator=g) * 2 - 4)
        t[:, :, 5:] = torch.rand(a.batch, 3, a.classes, G, G, device=dev, generator=g)
        levels.append(t.view(a.batch, 3 * (5 + a.classes), G, G))
    lib = L.load()
    if a.layout == "channels_last":
        levels = [t.view(a.batch, 3, 5 + a.classes, t.shape[2], t.shape[3]).permute(0, 1, 3, 4, 2).contiguous() for t in levels]
    d = _yolo_desc(levels, 3, a.conf_thres, 0.4, None, None, None, a.layout)
    n, n_pad = ctypes.c_int32(), ctypes.c_int32()
    lib.b200det_yolo_num_candidates(ctypes.byref(d), ctypes.byref(n), ctypes.byref(n_pad))
    nb = lib.b200det_yolo_workspace_bytes(ctypes.byref(d))
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    rows = torch.empty(a.batch, n_pad.value, 7, device=dev)
    cnt = torch.empty(a.batch, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    stages = [
        ("reset", lambda: lib.b200det_yolo_stage_reset(ctypes.byref(d), ws.data_ptr(), nb, st)),
        ("decode", lambda: lib.b200det_yolo_stage_decode(ctypes.byref(d), ws.data_ptr(), nb, st)),
        ("sort", lambda: lib.b200det_yolo_stage_sort(ctypes.byref(d), ws.data_ptr(), nb, st)),
        ("nms", lambda: lib.b200det_yolo_stage_nms(ctypes.byref(d), ws.data_ptr(), nb, st)),
        ("emit", lambda: lib.b200det_yolo_stage_emit(ctypes.byref(d), ws.data_ptr(), nb, rows.data_ptr(), None, cnt.data_ptr(), st)),
    ]
    times = {k: [] for k, _ in stages}
    for it in range(a.iters + 3):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(stages) + 1)]
        evs[0].record()
        for i, (k, f) in enumerate(stages):
            rc = f()
            assert rc == 0, (k, rc, lib.b200det_last_error())
            evs[i + 1].record()
        torch.cuda.synchronize()
        if it >= 3:
            for i, (k, _) in enumerate(stages):
                times[k].append(evs[i].elapsed_time(evs[i + 1]) * 1e3)
    med = {k: sorted(v)[len(v) // 2] for k, v in times.items()}
    total = sum(med.values())
    head_bytes = a.batch * n.value * (5 + a.classes) * 4
    out = dict(config=vars(a), N=n.value, kept=cnt.cpu().tolist()[:4], us=med, total_us=total,
               img_per_s=a.batch / (total * 1e-6), decode_GBps=head_bytes / (med["decode"] * 1e-6) / 1e9,
               ws_MB=nb / 1e6)
    print(json.dumps(out))
    if a.trace_nms:
        tr = torch.zeros(a.batch * a.classes * 8, dtype=torch.int64, device=dev)
        lib.b200det_debug_set_nms_trace.argtypes = [ctypes.c_void_p]
        assert lib.b200det_debug_set_nms_trace(tr.data_ptr()) == 0
        stages[3][1]()
        torch.cuda.synchronize()
        lib.b200det_debug_set_nms_trace(None)
        t = tr.cpu().view(-1, 8)
        t = t[t[:, 0] > 0]
        t0 = t[:, 0].min()
        names = ["load", "phaseB", "sweep", "owners", "merge"]
        d = (t[:, 1:6] - t[:, 0:5]).double() / 1e3
        life = (t[:, 5] - t[:, 0]).double() / 1e3
        print(f"{t.shape[0]} CTAs; span {(t[:, 5].max() - t0).item() / 1e3:.1f} us; CTA life mean {life.mean():.1f} max {life.max():.1f} us")
        print(" ".join(f"{nm} {d[:, i].mean():5.2f}/{d[:, i].max():5.2f}" for i, nm in enumerate(names)))
    if a.trace_sort:
        per_cta = 6656                                    # clustersort.cu: kCsCap; cluster = 1, 2, 4, 8 or 16 CTAs per image
        CL = next(c for c in (1, 2, 4, 8, 16) if n_pad.value <= c * per_cta)
        tr = torch.zeros(a.batch * CL * 64, dtype=torch.int64, device=dev)
        lib.b200det_debug_set_trace.argtypes = [ctypes.c_void_p]
        assert lib.b200det_debug_set_trace(tr.data_ptr()) == 0
        stages[2][1]()
        torch.cuda.synchronize()
        lib.b200det_debug_set_trace(None)
        t = tr.cpu().view(-1, 8, 8)
        t = t[(t[:, 0, 0] > 0)]
        t0 = t[:, 0, 0].min()
        names = ["load", "rank", "scan+csync", "exchange", "scatter", "csync"]
        print(f"{t.shape[0]} CTAs traced; kernel span {(t[:, :, :7].max() - t0).item() / 1e3:.1f} us")
        st0 = ((t[:, 0, 0] - t0).double() / 1e3).sort().values
        print("CTA start times (us), deciles:", " ".join(f"{st0[int(q * (len(st0) - 1) / 10)].item():.1f}" for q in range(11)))
        for ps in range(6):
            if t[:, ps, 0].max() == 0:
                continue
            d = (t[:, ps, 1:7] - t[:, ps, 0:6]).double() / 1e3
            start = (t[:, ps, 0] - t0).double() / 1e3
            print(f"pass {ps}: start {start.min():6.1f}..{start.max():6.1f} us | " +
                  " ".join(f"{nm} {d[:, i].mean():5.2f}/{d[:, i].max():5.2f}" for i, nm in enumerate(names)))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — benchmark of the detection post-processing / target-assignment hot path.

Headline metric (BASELINE.json): decode+NMS images/sec at batch 64, 640x640, YOLOv5s COCO (C=80) heads.
A "step" is one pass of the hot path over one batch of synthetic input per GPU (weak scaling: the per-rank batch is
fixed, ranks own disjoint image blocks, no data-path collective).

  python bench.py [--config NAME] [--gpus N] [--steps K] [--warmup W]        this repo's CUDA path
  python bench.py --impl reference [--config NAME] [--steps K] [--warmup W]  the UNMODIFIED reference on the host cores
  torchrun --nproc-per-node N ... bench.py --gpus N ...                      one rank per GPU

--config (default `headline`; every BASELINE.json configuration is a bench line):
  headline   YOLOv5s 640x640 COCO heads, batch 64, decode + filter + sort + class-aware merge-NMS      (the BASELINE metric)
  cfg1       YOLOv5s 640x640 VOC (20 classes), batch 1                                                 (BASELINE configs[0])
  cfg2       YOLOv3 416x416 COCO, batch 64                                                             (configs[1])
  ssd300     SSD300 (8 732 priors) prior decode + top-100 greedy NMS, batch 32                         (configs[2])
  retina800  RetinaNet 800x800 (120 087 anchors), batch 32                                             (configs[2])
  cfg4       YOLOv5s training-step build_targets_v5 + fused GIoU / objectness / class loss fwd+bwd, batch 64   (configs[3])
  crowd512   dense-crowd YOLOv5l 1280x1280, 5 classes, conf_thres 0.001, 64 images per GPU (512 over 8 GPUs) followed by the
             NCCL all-gather of the detections, timed as its own leg (`detection_allgather_ms`)        (configs[4])

Prints ONE JSON line (rank 0):
  value            whole-job images/s through the PUBLIC API (`od.non_max_suppression(None, levels)` etc. on device-resident
                   inputs; CUDA events around the wrapper call, its host sync on the counts and the list build included —
                   SURVEY 8d)
  pipeline_value   the same work enqueued through the C-ABI stage entry points without any host sync (round 1's `value`)
  e2e              the same metric through the public host-input API: pinned HOST buffers in, H2D and D2H inside the timed
                   region
  roofline         achieved HBM GB/s of the dominant streaming kernel (CUDA events around that launch, inside the timed
                   region) against the measured copy peak
  cpu_baseline     the reference's own CPU code (oracle/_ref, the unmodified reference staged by oracle/stage_ref.py) timed on
                   this host on a bounded sample
"""
from __future__ import annotations

import argparse
import ctypes
import gc
import json
import os
import statistics
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch

L2_BYTES = 126e6

CONFIGS = {
    "headline": dict(kind="yolo", name="yolov5s_640_coco_bs64", model="yolov5", img=640, classes=80, anchors=3, batch=64, seed=1234,
                     metric="decode+NMS images/sec (bs64, 640x640, YOLOv5s COCO heads)"),
    "cfg1": dict(kind="yolo", name="yolov5s_640_voc_bs1", model="yolov5", img=640, classes=20, anchors=3, batch=1, seed=1,
                 metric="decode+NMS images/sec (bs1, 640x640, YOLOv5s VOC heads)"),
    "cfg2": dict(kind="yolo", name="yolov3_416_coco_bs64", model="yolov3", img=416, classes=80, anchors=3, batch=64, seed=2,
                 metric="decode+NMS images/sec (bs64, 416x416, YOLOv3 COCO heads)"),
    "ssd300": dict(kind="prior", name="ssd300_coco_bs32", which="SSD", classes=80, batch=32, seed=3,
                   metric="prior decode + top-100 NMS images/sec (bs32, SSD300, 8732 priors, 80 classes)"),
    "retina800": dict(kind="prior", name="retinanet800_coco_bs32", which="RetinaNet", classes=80, batch=32, seed=3,
                      metric="prior decode + top-100 NMS images/sec (bs32, RetinaNet 800x800, 120087 anchors, 80 classes)"),
    "cfg4": dict(kind="targets", name="yolov5s_640_targets_bs64", img=640, classes=80, batch=64, seed=4,
                 metric="build_targets_v5 + GIoU/obj/cls loss fwd+bwd images/sec (bs64, 640x640, YOLOv5s COCO heads, <=100 boxes/image)"),
    "crowd512": dict(kind="yolo", name="yolov5l_1280_crowd_64_per_gpu", model="yolov5", img=1280, classes=5, anchors=3, batch=64,
                     seed=5, crowd=True, conf_thres=0.001,
                     metric="decode+NMS images/sec (dense crowd, 1280x1280 YOLOv5l heads, 5 classes, conf_thres 0.001, 64 images/GPU)"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default: 500; fewer for the slow configs)")
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="headline", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=None, help="images per GPU (default: the configuration's)")
    ap.add_argument("--e2e-steps", type=int, default=40)
    ap.add_argument("--gather-iters", type=int, default=20)
    ap.add_argument("--cpu-budget-s", type=float, default=150.0, help="wall budget of the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


# ------------------------------------------------------------------------------------------------------------------
# plumbing: clocks, affinity, stdout
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.ok = False
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover - depends on the box
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.01)          # ~100 samples per second: enough for the median, and the timed host loop keeps the GIL

    def __enter__(self):
        if self.ok:
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
            t0 = time.time()
            while not self.samples and time.time() - t0 < 2.0:      # the first NVML query can take a while
                time.sleep(0.001)
        return self

    def __exit__(self, *a):
        if self.ok:
            self._stop.set()
            self.t.join()

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def bind_to_gpu_numa_node(index):
    """Pin this rank to the CPUs NVML reports as local to its GPU before any pinned host buffer is allocated, so that the
    H2D / D2H traffic of the `e2e` leg does not cross the socket interconnect (8 ranks share one host)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


_JSON_FD = None


def claim_stdout():
    """Keep fd 1 for the ONE JSON line: libraries that write to the C stdout (NCCL prints its version there) are
    pointed at stderr for the rest of the run."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit_json(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


# ------------------------------------------------------------------------------------------------------------------
# synthetic inputs (CPU, seeded)
# ------------------------------------------------------------------------------------------------------------------
def yolo_inputs(cfg, batch, seed):
    from objectdetectionpl_b200 import synth
    if cfg.get("crowd"):
        grids = [cfg["img"] // 8, cfg["img"] // 16, cfg["img"] // 32]
        return synth.yolo_crowd(batch, cfg["anchors"], cfg["classes"], grids, cfg["img"], seed=seed), grids
    grids = synth.grids_for(cfg["model"], cfg["img"])
    return synth.yolo_planar(batch, cfg["anchors"], cfg["classes"], grids, cfg["img"], seed, v5_view=(cfg["model"] == "yolov5")), grids


def prior_inputs(cfg, batch, seed):
    from objectdetectionpl_b200 import synth
    pri = synth.ssd_priors() if cfg["which"] == "SSD" else synth.retina_priors(800)
    loc, cls = synth.prior_heads(batch, pri.shape[0], cfg["classes"], seed)
    return pri, loc, cls


def targets_inputs(cfg, batch, seed):
    from objectdetectionpl_b200 import synth
    C, img = cfg["classes"], cfg["img"]
    g = torch.Generator().manual_seed(seed + 100)
    heads = [torch.randn(batch, 3, img // s, img // s, 5 + C, generator=g) for s in (8, 16, 32)]
    tg = synth.labels(batch, C, seed)
    return heads, tg


def base_config(cfg, batch, world):
    c = {"workload": cfg["name"], "images_per_gpu": batch, "global_batch": batch * world, "classes": cfg["classes"]}
    if cfg["kind"] == "yolo":
        from objectdetectionpl_b200 import synth
        grids = [cfg["img"] // 8, cfg["img"] // 16, cfg["img"] // 32] if cfg.get("crowd") else synth.grids_for(cfg["model"], cfg["img"])
        c["candidates_per_image"] = sum(cfg["anchors"] * g * g for g in grids)
        c["conf_thres"] = cfg.get("conf_thres", "reference-forced -0.0151 (all candidates survive)")
        c["nms_thres"] = 0.4
    elif cfg["kind"] == "prior":
        c["priors_per_image"] = 8732 if cfg["which"] == "SSD" else 120087
        c["topk"], c["class_thresh"], c["nms_thresh"] = 100, 0.45, 0.5
    else:
        c["max_boxes_per_image"] = 100
    return c


# ------------------------------------------------------------------------------------------------------------------
# the reference arm / CPU baseline: the reference's OWN code on the host cores
# ------------------------------------------------------------------------------------------------------------------
def reference_runner(cfg, batch, seed):
    """Returns (step_fn, images_per_step, kind, what): one call of step_fn(i) runs a bounded sample of the workload through the
    reference's own CPU implementation (unmodified source from oracle/_ref resp. /root/reference)."""
    from oracle import ref_harness as rh           # checker / baseline only — never on the product path
    torch.set_num_threads(os.cpu_count() or 1)
    staged = "oracle/_ref (unmodified reference, staged by oracle/stage_ref.py)" if rh.is_staged_copy() else rh.REF_ROOT
    if not rh.available():
        return port_runner(cfg, batch, seed)
    if cfg["kind"] == "yolo":
        nimg = min(batch, 8)
        levels, _ = yolo_inputs(cfg, nimg, seed)
        N = base_config(cfg, batch, 1)["candidates_per_image"]
        if cfg.get("crowd"):
            # the unmodified reference overwrites its conf_thres argument (model/YOLOV5.py:164: conf_thres = -0.0151), so it cannot
            # run this configuration (conf_thres 0.001) at all; the line-by-line port honours the argument
            from oracle import ref_port as rp

            def step(i):
                rp.yolo_nms([t[i % nimg:i % nimg + 1] for t in levels], num_anchors=cfg["anchors"], conf_thres=cfg["conf_thres"],
                            compat=False)
            return step, 1, "port", (f"1 image per step ({N} candidates, ~10 % above conf_thres 0.001); oracle/ref_port.py::yolo_nms — the "
                                     "unmodified reference forces conf_thres = -0.0151 and cannot run this configuration")
        fn = rh.yolo_nms({"yolov5": 5, "yolov3": 3}[cfg["model"]])

        def step(i):
            with rh.cpu_only():
                fn(None, [t[i % nimg:i % nimg + 1].clone() for t in levels])
        return step, 1, "reference", (f"1 image per step ({N} candidates, all survive the forced conf_thres); "
                                      f"{fn.__qualname__} from {staged}; the loop is O(K*N) per image")
    if cfg["kind"] == "prior":
        pri, loc, cls = prior_inputs(cfg, batch, seed)
        fn = rh.ssd_nms(cfg["which"])
        me = types.SimpleNamespace(iou_boxes=pri)

        def step(i):
            with rh.cpu_only():
                fn(me, (loc, cls))
        return step, batch, "reference", f"the full batch of {batch} images per step; {fn.__qualname__} from {staged}"
    from objectdetectionpl_b200 import synth
    heads, tg = targets_inputs(cfg, batch, seed)
    with rh.cpu_only():
        crit = rh.losses().MultiScaleRegionLoss_v5(synth.YOLOV5_ANCHORS, None, None, None, None, cfg["classes"], cfg["img"])

    def step(i):
        with rh.cpu_only():
            p = [h.detach().requires_grad_(True) for h in heads]
            crit(p, tg.clone())["loss"].sum().backward()
    return step, batch, "reference", (f"the full batch of {batch} images per step; MultiScaleRegionLoss_v5 forward + backward "
                                      f"(LightningFunc/losses.py:98-152) from {staged}")


def port_runner(cfg, batch, seed):
    """Fallback when no reference tree is on the box (oracle/_ref is staged by build() where /root/reference exists): the
    line-by-line port of the same functions, oracle/ref_port.py (`kind: "port"`)."""
    from oracle import ref_port as rp
    why = "no reference tree on this box: oracle/ref_port.py, the line-by-line port"
    if cfg["kind"] == "yolo":
        nimg = min(batch, 8)
        levels, _ = yolo_inputs(cfg, nimg, seed)
        kw = dict(conf_thres=cfg["conf_thres"], compat=False) if "conf_thres" in cfg else {}
        return (lambda i: rp.yolo_nms([t[i % nimg:i % nimg + 1] for t in levels], num_anchors=cfg["anchors"], **kw)), 1, "port", \
            f"1 image per step; {why}"
    if cfg["kind"] == "prior":
        pri, loc, cls = prior_inputs(cfg, batch, seed)
        return (lambda i: rp.ssd_nms(loc, cls, pri)), batch, "port", f"the full batch of {batch} images per step; {why}"
    from objectdetectionpl_b200 import synth
    heads, tg = targets_inputs(cfg, batch, seed)
    stride = torch.tensor([8., 16., 32.])
    anchors = torch.tensor(synth.YOLOV5_ANCHORS).float().view(3, -1, 2) / stride.view(-1, 1, 1)

    def step(i):
        p = [h.detach().requires_grad_(True) for h in heads]
        tcls, tbox, idx, anch = rp.build_targets_v5([t.shape for t in p], tg, anchors, 3, 3)
        l = 0
        for k in range(3):
            giou, _ = rp.v5_match_level(p[k], tbox[k], idx[k], anch[k])
            l = l + (1.0 - giou).mean()
        l.backward()
    return step, batch, "port", f"the full batch of {batch} images per step (build_targets_v5 + matched-row GIoU fwd+bwd only); {why}"


def run_cpu(cfg, batch, seed, steps, warmup, budget_s):
    """-> dict(value images/s, steps done, warmup, cores, kind, sample)"""
    step, per_step, kind, what = reference_runner(cfg, batch, seed)
    t_start = time.perf_counter()
    times = []
    w_eff = 0
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step(i)
        dt = time.perf_counter() - t0
        if i < warmup and (time.perf_counter() - t_start) + 2 * dt < 0.3 * budget_s:
            w_eff += 1
            continue
        times.append(dt)
        if (time.perf_counter() - t_start) + dt > budget_s:
            break
    per = sum(times) / len(times)
    return dict(value=per_step / per, done=len(times), warmup=w_eff, cores=os.cpu_count() or 1, kind=kind,
                sample=f"{len(times)} timed step(s) after {w_eff} warm-up, {what}; all host threads "
                       f"(torch.set_num_threads({os.cpu_count()})); bounded by a {budget_s:.0f}s wall budget")


def run_reference(args, cfg, rank):
    if rank != 0:
        return
    batch = args.batch or cfg["batch"]
    r = run_cpu(cfg, batch, cfg["seed"], args.steps or 3, min(args.warmup, 1), args.cpu_budget_s)
    line = {
        "impl": "reference", "metric": cfg["metric"], "value": r["value"], "unit": "images/s", "n_gpus": args.gpus, "steps": r["done"],
        "warmup": r["warmup"], "ms_per_step": 1e3 / r["value"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": base_config(cfg, batch, 1),
        "cpu_baseline": {"value": r["value"], "unit": "images/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_json(line)


# ------------------------------------------------------------------------------------------------------------------
# workloads on the GPU
# ------------------------------------------------------------------------------------------------------------------
def rotation(nbytes):
    """Input sets to rotate through so that consecutive steps never find their input in L2 (126 MB)."""
    if nbytes > 1.5 * L2_BYTES:
        return 1
    return max(2, min(128, int(2.5 * L2_BYTES // max(nbytes, 1)) + 1))


class YoloWorkload:
    def __init__(self, cfg, batch, seed, dev):
        import objectdetectionpl_b200 as od
        from objectdetectionpl_b200 import _lib as L
        from objectdetectionpl_b200.postprocess import _yolo_desc
        self.od, self.L, self.lib, self.cfg, self.dev, self.B = od, L, L.load(), cfg, dev, batch
        self.levels_cpu, self.grids = yolo_inputs(cfg, batch, seed)
        self.pinned = [t.pin_memory() for t in self.levels_cpu]
        self.A, self.C = cfg["anchors"], cfg["classes"]
        self.N = sum(self.A * g * g for g in self.grids)
        self.head_bytes = batch * self.N * (5 + self.C) * 4
        self.R = rotation(self.head_bytes)
        self.sets = [[t.to(dev) for t in self.pinned]]
        for r in range(1, self.R):                 # identical bytes at different addresses: results identical, L2 never warm
            self.sets.append([t.clone() for t in self.sets[0]])
        self.compat = "conf_thres" not in cfg
        self.thr = od.YOLO_FORCED_CONF_THRES if self.compat else cfg["conf_thres"]
        self.kw = {} if self.compat else dict(compat=False, conf_thres=cfg["conf_thres"])
        lib = self.lib
        self.descs = [_yolo_desc(s, self.A, self.thr, 0.4, None, None, None) for s in self.sets]
        n, n_pad = ctypes.c_int32(), ctypes.c_int32()
        L.check(lib.b200det_yolo_num_candidates(ctypes.byref(self.descs[0]), ctypes.byref(n), ctypes.byref(n_pad)))
        self.n_pad = n_pad.value
        self.ws_bytes = lib.b200det_yolo_workspace_bytes(ctypes.byref(self.descs[0]))
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        self.rows = torch.empty((batch, self.n_pad, 7), dtype=torch.float32, device=dev)
        self.count = torch.empty((batch,), dtype=torch.int32, device=dev)
        self.st = torch.cuda.current_stream(dev).cuda_stream
        self.kernel = "yolo_decode_filter_kernel<MODE=0, U=8, MINB=6> (fused decode + filter + class argmax + ordered compaction)"
        self.stage_names = ["decode", "sort", "nms", "emit"]
        # our kernels per step: K1, sort, NMS, emit (+ the emit prefix in the packed public-API form).  The sort is ONE cluster
        # launch up to 53 248 slots per image; larger images take the dense route: compaction, cluster of 4, and the gated
        # multi-launch sort (histogram, class offsets, 4 score passes, class pass — empty launches unless an image overflows)
        sort_launches = 1 if self.n_pad <= 53248 else 9
        self.launches_pipe = 3 + sort_launches
        self.launches_api = 4 + sort_launches
        self.alg_bytes = self.head_bytes
        self.l2 = (f"inputs {self.head_bytes / 1e6:.1f} MB per GPU and step > 126 MB L2 (no flush needed)" if self.R == 1 else
                   f"{self.R} input sets of {self.head_bytes / 1e6:.1f} MB rotated ({self.R * self.head_bytes / 1e6:.0f} MB > 2x the 126 MB L2)")

    def api_step(self, i):
        out = self.od.non_max_suppression(None, self.sets[i % self.R], **self.kw)
        return out

    def pipe_step(self, i, ev=None, split=False):
        L, lib, st = self.L, self.lib, self.st
        dref, wp, wb = ctypes.byref(self.descs[i % self.R]), self.ws.data_ptr(), self.ws_bytes
        L.check(lib.b200det_yolo_stage_reset(dref, wp, wb, st))
        if ev is not None:
            ev[0].record()
        L.check(lib.b200det_yolo_stage_decode(dref, wp, wb, st))                 # K1 alone between ev[0] and ev[1]
        if ev is not None:
            ev[1].record()
        L.check(lib.b200det_yolo_stage_sort(dref, wp, wb, st))
        if split:
            ev[2].record()
        L.check(lib.b200det_yolo_stage_nms(dref, wp, wb, st))
        if split:
            ev[3].record()
        L.check(lib.b200det_yolo_stage_emit(dref, wp, wb, self.rows.data_ptr(), None, self.count.data_ptr(), st))
        if split:
            ev[4].record()

    def kept(self):
        return int(self.count.sum().item())

    def e2e_run(self, n):
        """`non_max_suppression_host_async`: 8-image chunks, H2D of chunk k+1 / pipeline of chunk k / D2H of chunk k-1 overlap;
        batch i+1 is submitted before batch i is collected.  A step is collected when its detections are in pinned host
        memory and the host has read the row counts."""
        od, prev, nrows = self.od, None, 0
        for _ in range(n):
            h = od.non_max_suppression_host_async(None, self.pinned, device=self.dev, **self.kw)
            if prev is not None:
                nrows = sum(x.shape[0] for x in prev.result() if x is not None)
            prev = h
        self.e2e_rows = sum(x.shape[0] for x in prev.result() if x is not None)
        return self.e2e_rows

    def e2e_bytes(self):
        # device -> host: the kept rows (the emit kernel writes them, packed, straight into pinned host memory) + per 8-image
        # chunk the count | offsets words
        chunks = (self.B + 7) // 8
        return sum(t.numel() * 4 for t in self.pinned), self.e2e_rows * 28 + chunks * (2 * 8 + 1) * 4

    e2e_api = ("objectdetectionpl_b200.non_max_suppression_host_async (pinned host in, 8-image chunks, H2D on one stream and the "
               "pipeline on another; the emit kernel writes the kept rows straight into pinned host memory — no D2H copy; batch "
               "i+1 submitted before batch i is collected)")


class PriorWorkload:
    def __init__(self, cfg, batch, seed, dev):
        import objectdetectionpl_b200 as od
        from objectdetectionpl_b200 import _lib as L
        self.od, self.L, self.lib, self.cfg, self.dev, self.B = od, L, L.load(), cfg, dev, batch
        pri, loc, cls = prior_inputs(cfg, batch, seed)
        self.P, self.C = pri.shape[0], cfg["classes"]
        self.pri = pri.to(dev)
        self.pinned = [loc.pin_memory(), cls.pin_memory()]
        self.head_bytes = (loc.numel() + cls.numel()) * 4
        self.R = rotation(self.head_bytes)
        self.sets = [[t.to(dev) for t in self.pinned]]
        for r in range(1, self.R):
            self.sets.append([t.clone() for t in self.sets[0]])
        self.me = types.SimpleNamespace(iou_boxes=self.pri)
        self.descs = []
        for loc_d, cls_d in self.sets:
            d = L.PriorDesc()
            d.batch, d.num_priors, d.num_classes = batch, self.P, self.C
            d.loc, d.cls, d.priors = loc_d.data_ptr(), cls_d.data_ptr(), self.pri.data_ptr()
            d.topk, d.nms_thresh, d.class_thresh, d.mode_min, d.compat = 100, 0.5, 0.45, 0, 1
            self.descs.append(d)
        self.ws_bytes = self.lib.b200det_prior_workspace_bytes(ctypes.byref(self.descs[0]))
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        self.rows = torch.empty((batch, 100, 7), dtype=torch.float32, device=dev)
        self.count = torch.empty((2, batch), dtype=torch.int32, device=dev)
        self.st = torch.cuda.current_stream(dev).cuda_stream
        self.kernel = "prior_decode_filter_kernel<true,true> (prior decode + sigmoid-argmax + score filter + ordered compaction)"
        self.stage_names = ["decode", "select+nms"]
        self.launches_pipe = self.launches_api = 3            # K1', radix select, NMS (the counts copy is a cudaMemcpyAsync)
        self.alg_bytes = self.head_bytes
        self.l2 = (f"inputs {self.head_bytes / 1e6:.1f} MB per GPU and step > 126 MB L2 (no flush needed)" if self.R == 1 else
                   f"{self.R} input sets of {self.head_bytes / 1e6:.1f} MB rotated ({self.R * self.head_bytes / 1e6:.0f} MB > 2x the 126 MB L2)")

    def api_step(self, i):
        loc, cls = self.sets[i % self.R]
        return self.od.prior_non_max_suppression(self.me, (loc, cls))

    def pipe_step(self, i, ev=None, split=False):
        L, lib, st = self.L, self.lib, self.st
        dref, wp, wb = ctypes.byref(self.descs[i % self.R]), self.ws.data_ptr(), self.ws_bytes
        if ev is not None:
            ev[0].record()
        L.check(lib.b200det_prior_stage_decode(dref, wp, wb, st))                # K1' alone between ev[0] and ev[1]
        if ev is not None:
            ev[1].record()
        L.check(lib.b200det_prior_stage_select_nms(dref, wp, wb, self.rows.data_ptr(), None, self.count[0].data_ptr(),
                                                   self.count[1].data_ptr(), st))
        if split:
            ev[2].record()

    def kept(self):
        return int(self.count[0].sum().item())

    def e2e_run(self, n):
        nrows = 0
        for _ in range(n):
            out = self.od.prior_non_max_suppression_host(self.me, tuple(self.pinned), device=self.dev)
            nrows = sum(x.shape[0] for x in out)
        return nrows

    def e2e_bytes(self):
        return sum(t.numel() * 4 for t in self.pinned), self.B * 100 * 7 * 4 + 2 * self.B * 4

    e2e_api = "objectdetectionpl_b200.prior_non_max_suppression_host (pinned host loc/cls in, pinned host rows out)"


class TargetsWorkload:
    def __init__(self, cfg, batch, seed, dev):
        import objectdetectionpl_b200 as od
        from objectdetectionpl_b200 import synth
        self.od, self.cfg, self.dev, self.B, self.C = od, cfg, dev, batch, cfg["classes"]
        heads, tg = targets_inputs(cfg, batch, seed)
        self.pinned = [h.pin_memory() for h in heads]
        self.tg_pinned = tg.pin_memory()
        self.tg = tg.to(dev)
        self.nt = int(tg.shape[0])
        self.p = [h.to(dev).requires_grad_(True) for h in self.pinned]
        stride = torch.tensor([8., 16., 32.])
        self.anchors = (torch.tensor(synth.YOLOV5_ANCHORS).float().view(3, -1, 2) / stride.view(-1, 1, 1)).to(dev)
        self.head_bytes = sum(h.numel() * 4 for h in heads)
        self.R = 1
        tcls, _, _, _ = od.build_targets_v5([tuple(h.shape) for h in heads], self.tg, self.anchors, 3, 3)
        self.m = [int(t.shape[0]) for t in tcls]
        M, F = sum(self.m), 5 + self.C
        cells = sum(h.numel() // F for h in heads)
        # nt*24 in + sum m_i (4*8 + 16 + 8 + 8) out + gathered sum m_i (5+C)*4; plus what the fused loss tail must touch:
        # the objectness logit of every cell (read, and its gradient written) and the gradient rows of the matched cells
        self.alg_bytes = self.nt * 24 + M * (32 + 16 + 8 + 8) + 2 * M * F * 4 + 2 * cells * 4
        self.kernel = ("whole step: build_targets_v5 + the fused v5 loss forward / backward, one launch per stage for all levels "
                       "(9 launches; the 548 MB gradient stream of the objectness backward is its largest kernel)")
        self.stage_names = []
        # build_targets_v5 (1); forward, all levels per launch: match, tobj scatter, matched-row terms, objectness terms,
        # means + combination (5); backward: combination (1), objectness gradient stream, matched-row gradients (2)
        self.launches_pipe = self.launches_api = 1 + 5 + 1 + 2
        self.l2 = f"heads {self.head_bytes / 1e6:.1f} MB per GPU > 126 MB L2 (no flush needed)"

    def api_step(self, i):
        for t in self.p:
            t.grad = None
        m = self.od.v5_loss(self.p, self.tg, self.anchors, 3, 3, self.C)
        m["loss"].backward()
        return m

    def pipe_step(self, i, ev=None, split=False):
        if ev is not None:
            ev[0].record()
        self.api_step(i)
        if ev is not None:
            ev[1].record()

    def kept(self):
        return sum(self.m)

    def e2e_run(self, n):
        dev, val = self.dev, 0.0
        for _ in range(n):
            p = [h.to(dev, non_blocking=True).requires_grad_(True) for h in self.pinned]
            tg = self.tg_pinned.to(dev, non_blocking=True)
            m = self.od.v5_loss(p, tg, self.anchors, 3, 3, self.C)
            m["loss"].backward()
            val = float(torch.cat([m["loss"], m["Localization"], m["Classification"], m["Conf_obj"]]).cpu()[0])
        return val

    def e2e_bytes(self):
        return sum(t.numel() * 4 for t in self.pinned) + self.tg_pinned.numel() * 4, 16

    e2e_api = ("pinned host heads + labels -> device, objectdetectionpl_b200.v5_loss(...)['loss'].backward(), the four loss scalars "
               "read back (the head gradients stay on the device, where the backbone's backward consumes them)")


def main():
    args = parse()
    cfg = CONFIGS[args.config]
    rank, world, local = dist_env()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return

    import torch.distributed as dist
    import objectdetectionpl_b200 as od

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    local_cpus = bind_to_gpu_numa_node(local)      # before any pinned allocation: keeps the host buffers on the GPU's NUMA node
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch or cfg["batch"]
    K = args.steps or (500 if cfg["kind"] != "targets" else 200)
    W = max(args.warmup, 3)
    wl = {"yolo": YoloWorkload, "prior": PriorWorkload, "targets": TargetsWorkload}[cfg["kind"]](cfg, B, cfg["seed"] + rank, dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for i in range(W):
        wl.api_step(i)
        wl.pipe_step(i)
    torch.cuda.synchronize(dev)

    # ---- timed region: (A) the public API, (B) the stage entry points with CUDA events around the streaming kernel -----
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k1_ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(K)]
    barrier()
    gc.collect()
    gc.disable()            # as `timeit` does: no collector pauses inside the timed host loops
    try:
        with ClockSampler(local) as clk:
            a0.record()
            for i in range(K):
                wl.api_step(i)
            a1.record()
            barrier()
            b0.record()
            for i in range(K):
                wl.pipe_step(i, k1_ev[i])
            b1.record()
            barrier()
    finally:
        gc.enable()
    api_ms = max_over_ranks(a0.elapsed_time(a1))
    pipe_ms = max_over_ranks(b0.elapsed_time(b1))
    value = world * B * K / (api_ms * 1e-3)
    pipeline_value = world * B * K / (pipe_ms * 1e-3)
    k1_us = statistics.mean(k1_ev[i][0].elapsed_time(k1_ev[i][1]) for i in range(K)) * 1e3
    kept_total = wl.kept()

    stage_us = {}
    if wl.stage_names:
        Ks, ns = min(K, 50), len(wl.stage_names)
        sev = [[torch.cuda.Event(enable_timing=True) for _ in range(ns + 1)] for _ in range(Ks)]
        for i in range(Ks):
            wl.pipe_step(i, sev[i], split=True)
        torch.cuda.synchronize(dev)
        stage_us = {nm: statistics.mean(sev[i][k].elapsed_time(sev[i][k + 1]) for i in range(Ks)) * 1e3
                    for k, nm in enumerate(wl.stage_names)}

    # ---- end to end from pinned host buffers ---------------------------------------------------------------------------
    wl.e2e_run(3)
    Ke = max(1, min(args.e2e_steps, K))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    wl.e2e_run(Ke)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = world * B * Ke / (e2e_ms * 1e-3)
    h2d, d2h = wl.e2e_bytes()

    # ---- the one exchange step of the path: all-gather of the detections for mAP, timed as its own leg ---------------------
    gather = None
    if cfg["kind"] == "yolo" and (world > 1 or args.config == "crowd512"):
        rows, _, count = od.yolo_nms_raw(wl.sets[0], wl.A, wl.thr)
        send = torch.empty((B * rows.shape[1], 8), dtype=torch.float32, device=dev)
        for _ in range(3):
            g = od.dist.gather_detections_raw(rows, count, rank * B, send=send)
        barrier()
        gt = []
        for _ in range(max(1, args.gather_iters)):
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if world > 1:
                dist.barrier()
            g0.record()
            g = od.dist.gather_detections_raw(rows, count, rank * B, send=send)
            g1.record()
            torch.cuda.synchronize(dev)
            gt.append(g0.elapsed_time(g1))
        g_ms = max_over_ranks(statistics.median(gt))
        per = g.per_image()
        assert len(per) == world * B and sum(g.totals) == sum(0 if p is None else p.shape[0] for p in per)
        ids = g.rows[world - 1, :g.totals[world - 1], 7]
        assert int(ids.max().item()) == world * B - 1 and int(g.rows[0, 0, 7].item()) == 0
        payload = sum(g.totals) * 32
        gather = {"ms": g_ms, "rows_total": sum(g.totals), "payload_MB": payload / 1e6, "algbw_GBps": payload / (g_ms * 1e-3) / 1e9,
                  "how": "device-side pack kernel + all_gather_into_tensor(counts) + one host read + all_gather_into_tensor(rows) into a "
                         "persistent buffer; median of %d warm iterations, CUDA events, max over ranks" % max(1, args.gather_iters)}

    if rank == 0:
        peak, peak_src = measured_peak()
        # plain pinned-copy ceiling of this kind of box at this many ranks (tools/h2d_bandwidth.py under torchrun, all ranks copying
        # at once; profiles/h2d_ceiling.json), what the e2e leg is judged against
        ceiling = None
        hp = os.path.join(ROOT, "profiles", "h2d_ceiling.json")
        if os.path.exists(hp):
            try:
                ceiling = json.load(open(hp)).get(str(world))
            except Exception:
                ceiling = None
        achieved = wl.alg_bytes / (k1_us * 1e-6) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "k1_traffic.json")   # dram__bytes_read+write of one launch (ncu --set full), per config
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                traffic = float(tj.get(args.config, {}).get("dram_bytes_per_launch")) if args.config in tj else None
            except Exception:
                traffic = None
        c = base_config(cfg, B, world)
        c["parallelism"] = f"image-sharded x{world}, no data-path collective"
        c["l2"] = wl.l2
        line = {
            "metric": cfg["metric"], "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": api_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": c, "clocks": clk.summary(),
            "value_api": "public API on device-resident inputs, CUDA events around the wrapper calls (host sync on the counts and "
                         "the list build included)",
            "pipeline_value": pipeline_value, "pipeline_ms_per_step": pipe_ms / K,
            "api_over_pipeline": (api_ms / pipe_ms),
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke,
                    "ms_per_step": e2e_ms / Ke, "h2d_GBps": world * h2d / (e2e_ms / Ke * 1e-3) / 1e9,
                    "h2d_ceiling_gbs": ceiling["aggregate_GBps"] if ceiling else None,
                    "frac_of_h2d_ceiling": (world * h2d / (e2e_ms / Ke * 1e-3) / 1e9) / ceiling["aggregate_GBps"] if ceiling else None,
                    "api": wl.e2e_api},
            "gpu_launches": (wl.launches_api + wl.launches_pipe) * K,
            "gpu_launches_note": f"{wl.launches_api} of this repo's kernels per public-API step + {wl.launches_pipe} per stage-entry step, {K} steps each",
            "roofline": {"bound": "hbm", "kernel": wl.kernel + " — one launch per step, CUDA events around it in the timed region",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": wl.alg_bytes,
                         "pipeline_frac": (wl.alg_bytes + (kept_total * 28 if cfg["kind"] != "targets" else 0)) / (pipe_ms / K * 1e-3) / 1e9 / peak},
            "stages_us": stage_us, "k1_us_in_timed_region": k1_us, "rank_cpu_affinity": local_cpus,
            "kept_per_image_mean": kept_total / B, "us_per_image": api_ms / K * 1e3 / B,
        }
        if cfg["kind"] == "targets":
            line["targets_per_s"] = wl.nt * K / (api_ms * 1e-3) * world
            line["us_per_step"] = api_ms / K * 1e3
            line["matched_rows_per_level"] = wl.m
        if gather is not None:
            line["detection_allgather_ms"] = gather["ms"]
            line["detection_allgather"] = gather
        if world == 1 and not args.no_cpu_baseline:
            r = run_cpu(cfg, B, cfg["seed"], 2, 0, 25.0)
            line["cpu_baseline"] = {"value": r["value"], "unit": "images/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
        emit_json(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

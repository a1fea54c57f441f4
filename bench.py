#!/usr/bin/env python
"""bench.py — headline benchmark of the detection post-processing hot path.

Metric (BASELINE.json): decode+NMS images/sec at batch 64, 640x640, YOLOv5s COCO (C=80) heads.
A "step" is one pass of the hot path (fused decode+filter -> per-image sort -> class-aware merge-NMS
-> ordered emit) over one batch of 64 synthetic images per GPU (weak scaling: 64 images per rank).

  python bench.py [--gpus N] [--steps K] [--warmup W]                  this repo's CUDA path
  python bench.py --impl reference [--steps K] [--warmup W]            the reference algorithm on host cores
  torchrun --nproc-per-node N ... bench.py --gpus N ...                one rank per GPU, no data-path collective

Prints ONE JSON line (rank 0).  `value` = whole-job images/s with inputs resident in HBM; `e2e` = the same
metric through the public API from pinned HOST buffers (H2D of the heads and D2H of the detections inside
the timed region); `roofline` = achieved HBM GB/s of the dominant streaming kernel (fused decode+filter)
against the measured copy peak; `cpu_baseline` = the oracle port of the reference timed on this host.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch

WORKLOAD = dict(name="yolov5s_640_coco_bs64", model="yolov5", img=640, classes=80, anchors=3, batch=64, seed=1234)
METRIC = "decode+NMS images/sec (bs64, 640x640, YOLOv5s COCO heads)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=WORKLOAD["batch"], help="images per GPU (default 64)")
    ap.add_argument("--e2e-steps", type=int, default=50)
    ap.add_argument("--cpu-budget-s", type=float, default=150.0, help="wall budget of the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def make_inputs(batch, seed):
    from objectdetectionpl_b200 import synth
    w = WORKLOAD
    grids = synth.grids_for(w["model"], w["img"])
    return synth.yolo_planar(batch, w["anchors"], w["classes"], grids, w["img"], seed, v5_view=True), grids


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
            time.sleep(0.002)

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


def cpu_reference_run(levels_cpu, steps, warmup, budget_s):
    """The reference algorithm (oracle port of model/YOLOV5.py:157-218, torch eager on the host cores), one image
    per step.  O(K*N) per image (~10-20 s), so the number of steps is capped by a wall budget."""
    from oracle import ref_port as rp  # checker / baseline only
    torch.set_num_threads(os.cpu_count() or 1)
    B = levels_cpu[0].shape[0]
    times = []
    done = 0
    t_start = time.perf_counter()
    w_eff = min(warmup, 1)
    for i in range(w_eff + steps):
        img = [t[i % B:i % B + 1] for t in levels_cpu]
        t0 = time.perf_counter()
        rp.yolo_nms(img, num_anchors=WORKLOAD["anchors"])
        dt = time.perf_counter() - t0
        if i >= w_eff:
            times.append(dt)
            done += 1
        elapsed = time.perf_counter() - t_start
        if elapsed + dt > budget_s and done >= 1:
            break
    per_img = sum(times) / len(times)
    return 1.0 / per_img, done, w_eff


def run_reference(args, rank):
    if rank != 0:
        return
    levels, _ = make_inputs(min(args.batch, 8), WORKLOAD["seed"])
    ips, done, w_eff = cpu_reference_run(levels, args.steps, args.warmup, args.cpu_budget_s)
    cores = os.cpu_count() or 1
    sample = (f"{done} timed step(s) of 1 image each (25200 candidates, all survive) after {w_eff} warm-up; steps capped by a "
              f"{args.cpu_budget_s:.0f}s wall budget because the reference loop is O(K*N) per image")
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus, "steps": done,
        "warmup": w_eff, "ms_per_step": 1e3 / ips, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD["name"], "images_per_step": 1, "candidates_per_image": 25200, "classes": 80},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_json(line)


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


def main():
    args = parse()
    rank, world, local = dist_env()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import ctypes
    import torch.distributed as dist
    import objectdetectionpl_b200 as od
    from objectdetectionpl_b200 import _lib as L
    from objectdetectionpl_b200.postprocess import _yolo_desc

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    local_cpus = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    lib = L.load()
    w = WORKLOAD
    B = args.batch
    levels_cpu, grids = make_inputs(B, w["seed"] + rank)          # per-rank shard of the global batch (by image)
    pinned = [t.pin_memory() for t in levels_cpu]
    levels = [t.to(dev) for t in pinned]
    N = sum(w["anchors"] * g * g for g in grids)
    head_bytes = B * N * (5 + w["classes"]) * 4

    # ---- stage-wise descriptors (the pipeline call is exactly these four stage calls) -----------------------------
    d = _yolo_desc(levels, w["anchors"], od.YOLO_FORCED_CONF_THRES, 0.4, None, None, None)
    n, n_pad = ctypes.c_int32(), ctypes.c_int32()
    L.check(lib.b200det_yolo_num_candidates(ctypes.byref(d), ctypes.byref(n), ctypes.byref(n_pad)))
    ws_bytes = lib.b200det_yolo_workspace_bytes(ctypes.byref(d))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    rows = torch.empty((B, n_pad.value, 7), dtype=torch.float32, device=dev)
    count = torch.empty((B,), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    dref, wp, rp_, cp = ctypes.byref(d), ws.data_ptr(), rows.data_ptr(), count.data_ptr()

    def step(ev=None, split=False):
        L.check(lib.b200det_yolo_stage_reset(dref, wp, ws_bytes, st))       # zero-fill of the counter header
        if ev is not None:
            ev[0].record()
        L.check(lib.b200det_yolo_stage_decode(dref, wp, ws_bytes, st))      # K1 alone between ev[0] and ev[1]
        if ev is not None:
            ev[1].record()
        L.check(lib.b200det_yolo_stage_sort(dref, wp, ws_bytes, st))
        if split:
            ev[2].record()
        L.check(lib.b200det_yolo_stage_nms(dref, wp, ws_bytes, st))
        if split:
            ev[3].record()
        L.check(lib.b200det_yolo_stage_emit(dref, wp, ws_bytes, rp_, None, cp, st))
        if split:
            ev[4].record()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    K = args.steps
    # Inside the timed region only K1 is bracketed by events (roofline.achieved must be measured there); the split of
    # the other stages comes from a short extra pass afterwards so that its events do not sit in the headline loop.
    stage_ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(K)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with ClockSampler(local) as clk:
        t_begin.record()
        for i in range(K):
            step(stage_ev[i])
        t_end.record()
        barrier()
    total_ms = t_begin.elapsed_time(t_end)
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms_max = float(tmax.item())
    ms_per_step = total_ms_max / K
    value = world * B * K / (total_ms_max * 1e-3)
    k1_us = statistics.mean(stage_ev[i][0].elapsed_time(stage_ev[i][1]) for i in range(K)) * 1e3
    Ks = min(K, 50)
    split_ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(Ks)]
    for i in range(Ks):
        step(split_ev[i], split=True)
    torch.cuda.synchronize(dev)
    names = ["decode", "sort", "nms", "emit"]
    stage_us = {nm: statistics.mean(split_ev[i][k].elapsed_time(split_ev[i][k + 1]) for i in range(Ks)) * 1e3
                for k, nm in enumerate(names)}
    kept = count.cpu()
    kept_total = int(kept.sum())

    # ---- end-to-end through the public API from pinned host buffers -------------------------------------------------
    # non_max_suppression_host_async: chunks of 8 images, H2D of chunk k+1 / CUDA pipeline of chunk k / D2H of chunk k-1
    # overlap; batch i+1 is submitted before batch i is collected (two pinned result sets), so the host->device link stays
    # busy across steps.  A step is collected when its detections (padded rows + counts) are in pinned host memory and
    # the host has read them (row count per image).  Every step's H2D and D2H lie inside the timed region.
    def e2e_run(n):
        prev, nrows = None, 0
        for _ in range(n):
            h = od.non_max_suppression_host_async(None, pinned, device=dev)
            if prev is not None:
                nrows = sum(x.shape[0] for x in prev.result() if x is not None)
            prev = h
        return sum(x.shape[0] for x in prev.result() if x is not None)

    e2e_run(3)
    Ke = max(1, min(args.e2e_steps, K))
    barrier()
    t0 = time.perf_counter()
    e_begin, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_begin.record()
    nrows = e2e_run(Ke)
    e_end.record()
    barrier()
    e_ms = torch.tensor([e_begin.elapsed_time(e_end)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * B * Ke / (float(e_ms.item()) * 1e-3)
    h2d = sum(t.numel() * 4 for t in pinned)
    d2h = B * n_pad.value * 7 * 4 + B * 4          # padded rows [B, n_pad, 7] + counts (no host sync before the copy)

    # ---- the one exchange step of the path: all-gather of detections for mAP (outside the timed region) --------------
    gather_ms = None
    if world > 1:
        dets = od.non_max_suppression(None, levels)
        torch.cuda.synchronize(dev)
        g0 = time.perf_counter()
        allp = od.dist.gather_detections(dets, image_offset=rank * B, device=dev)
        torch.cuda.synchronize(dev)
        gather_ms = (time.perf_counter() - g0) * 1e3
        assert int(allp[:, 7].max().item()) == world * B - 1

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = head_bytes / (k1_us * 1e-6) / 1e9
        launches = 1 + 1 + 1 + 1                               # K1; cluster sort; NMS; emit (the reset stage launches nothing)
        traffic = None
        tp = os.path.join(ROOT, "profiles", "k1_traffic.json")   # dram__bytes_read+write of one K1 launch (ncu --set full)
        if os.path.exists(tp):
            try:
                traffic = float(json.load(open(tp))["dram_bytes_per_launch"])
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": w["name"], "images_per_gpu": B, "global_batch": B * world, "candidates_per_image": N,
                       "classes": w["classes"], "conf_thres": "reference-forced -0.0151 (all candidates survive)",
                       "nms_thres": 0.4, "parallelism": f"image-sharded x{world}, no data-path collective",
                       "l2": f"inputs {head_bytes / 1e6:.0f} MB per GPU > 126 MB L2 (no flush needed)"},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": Ke, "api": "objectdetectionpl_b200.non_max_suppression_host_async (pinned host in/out, 8-image chunks on 3 streams, batch i+1 submitted before batch i is collected)"},
            "gpu_launches": launches * K,
            "roofline": {"bound": "hbm", "kernel": "yolo_decode_filter_kernel<4,0,8,6> (one launch per step, CUDA events around it)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": head_bytes},
            "stages_us": stage_us, "k1_us_in_timed_region": k1_us, "rank_cpu_affinity": local_cpus, "kept_per_image_mean": kept_total / B, "workspace_mb": ws_bytes / 1e6,
        }
        if gather_ms is not None:
            line["detection_allgather_ms"] = gather_ms
        if world == 1 and not args.no_cpu_baseline:
            ips, done, w_eff = cpu_reference_run([t[:2] for t in levels_cpu], 1, 0, 60.0)
            line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": os.cpu_count() or 1, "kind": "port",
                                    "sample": f"{done} image of the same batch (25200 candidates, all survive), no warm-up, "
                                              "torch-CPU port of the reference loop with all host threads"}
        emit_json(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

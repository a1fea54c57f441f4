"""Detection post-processing with the reference's call signatures.

  non_max_suppression(self, predictions, conf_thres=0.5, nms_thres=0.4)            YOLOv3/v4/v5
      replaces model/YOLOV5.py:157, YOLOV3.py:273, YOLOV4.py:221
  non_max_suppression_v2(...)                                                       YOLOv2 (5 anchors)
      replaces model/YOLOV2.py:159
  prior_non_max_suppression(self, predictions, topk=100, nms_thresh=0.5, class_thresh=0.45, mode='union')
      replaces model/SSD.py:249 == model/RetinaNet.py:117
  decode_box(head, anchors, stride, mode)  — north-star API, restates the inline decode sites D1/D2.

Everything runs on the GPU through libb200det.so; CPU tensors raise.
"""
from __future__ import annotations

import collections
import ctypes
import threading
import time
from typing import List, Optional, Sequence

import torch

from . import _lib as L

YOLO_FORCED_CONF_THRES = -0.0151   # model/YOLOV5.py:164 — the reference overwrites its conf_thres argument

_DECODE = {None: L.DECODE_NONE, "none": L.DECODE_NONE, "yolo_exp": L.DECODE_YOLO_EXP, "yolov5": L.DECODE_YOLOV5,
           "yolov4_norm": L.DECODE_YOLOV4_NORM}


_LAYOUT = {None: L.LAYOUT_PLANAR, "planar": L.LAYOUT_PLANAR, "channels_last": L.LAYOUT_CHANNELS_LAST}


def _yolo_desc(levels: Sequence[torch.Tensor], num_anchors: int, conf_thres: float, nms_thres: float,
               decode, anchors, strides, layout=None, scale_x_y: float = 1.0) -> L.YoloDesc:
    if len(levels) == 0 or len(levels) > L.MAX_LEVELS:
        raise ValueError(f"need 1..{L.MAX_LEVELS} prediction levels, got {len(levels)}")
    d = L.YoloDesc()
    first = L.require_cuda(levels[0], "predictions[0]")
    B = first.shape[0]
    C = None
    for i, t in enumerate(levels):
        L.require_cuda(t, f"predictions[{i}]")
        if t.device != first.device:
            raise ValueError("all prediction levels must live on one device")
        if not t.is_contiguous():
            raise ValueError(f"predictions[{i}] must be contiguous (the reference .view()s it, model/YOLOV3.py:296)")
        if t.dim() < 4 or t.shape[0] != B:
            raise ValueError(f"predictions[{i}] has shape {tuple(t.shape)}; expected [B, A*(5+C), G, G] or [B, A, G, G, 5+C]")
        G = t.shape[2]                                   # grid_size = prediction.size(2), YOLOV3.py:290 / YOLOV5.py:175
        per = num_anchors * G * G * B
        if t.numel() % per:
            raise ValueError(f"predictions[{i}] of shape {tuple(t.shape)} is not [B, {num_anchors}, 5+C, {G}, {G}] storage")
        fields = t.numel() // per
        if C is None:
            C = fields - 5
        if fields - 5 != C or C < 1:
            raise ValueError("all levels must carry the same 5+C fields (C >= 1)")
        d.head[i] = t.data_ptr()
        d.grid[i] = G
    d.batch, d.num_anchors, d.num_classes, d.num_levels = B, num_anchors, C, len(levels)
    d.decode_mode = _DECODE[decode]
    d.scale_x_y = float(scale_x_y)
    if d.decode_mode != L.DECODE_NONE:
        if d.decode_mode == L.DECODE_YOLOV4_NORM and strides is None:
            strides = [1.0] * len(levels)                    # D3 normalises by the grid size; no stride
        if anchors is None or strides is None or len(anchors) != len(levels) or len(strides) != len(levels):
            raise ValueError("decode modes need per-level `anchors` ([A,2] each) and `strides`")
        for i in range(len(levels)):
            d.stride[i] = float(strides[i])
            a = torch.as_tensor(anchors[i], dtype=torch.float32).reshape(-1, 2).cpu()
            if a.shape[0] != num_anchors:
                raise ValueError(f"anchors[{i}] must hold {num_anchors} (w,h) pairs")
            for k in range(num_anchors):
                d.anchors[i][k][0] = float(a[k, 0])
                d.anchors[i][k][1] = float(a[k, 1])
    d.conf_thres, d.nms_thres = float(conf_thres), float(nms_thres)
    if layout not in _LAYOUT:
        raise ValueError(f"layout must be 'planar' or 'channels_last', got {layout!r}")
    d.layout = _LAYOUT[layout]
    if d.layout == L.LAYOUT_CHANNELS_LAST:
        for i, t in enumerate(levels):
            if t.dim() != 5 or t.shape[1] != num_anchors or t.shape[2] != t.shape[3] or t.shape[4] != C + 5:
                raise ValueError(f"layout='channels_last' needs predictions[{i}] of shape [B, {num_anchors}, G, G, 5+C], "
                                 f"got {tuple(t.shape)}")
    return d


def yolo_nms_raw(levels: Sequence[torch.Tensor], num_anchors: int = 3, conf_thres: float = YOLO_FORCED_CONF_THRES,
                 nms_thres: float = 0.4, decode=None, anchors=None, strides=None, want_index: bool = False, layout=None,
                 scale_x_y: float = 1.0):
    """Enqueue the whole pipeline; returns device tensors (rows [B,n_pad,7], index [B,n_pad]|None, count [B])
    without synchronising — the building block for benchmarks and CUDA-graph capture."""
    lib = L.load()
    d = _yolo_desc(levels, num_anchors, conf_thres, nms_thres, decode, anchors, strides, layout, scale_x_y)
    dev = levels[0].device
    n, n_pad = ctypes.c_int32(), ctypes.c_int32()
    L.check(lib.b200det_yolo_num_candidates(ctypes.byref(d), ctypes.byref(n), ctypes.byref(n_pad)), "yolo_num_candidates")
    ws_bytes = lib.b200det_yolo_workspace_bytes(ctypes.byref(d))
    with torch.cuda.device(dev):
        ws = L.workspace(ws_bytes, dev)
        rows = torch.empty((d.batch, n_pad.value, 7), dtype=torch.float32, device=dev)
        index = torch.empty((d.batch, n_pad.value), dtype=torch.int32, device=dev) if want_index else None
        count = torch.empty((d.batch,), dtype=torch.int32, device=dev)
        L.check(lib.b200det_yolo_nms(ctypes.byref(d), ws.data_ptr(), ws.numel(), rows.data_ptr(),
                                     index.data_ptr() if want_index else None, count.data_ptr(), L.stream_ptr(dev)),
                "yolo_nms")
    return rows, index, count


class _YoloPlan:
    """Everything about a `non_max_suppression` call that only depends on the head SHAPES (and the scalar options): the
    filled-in descriptor, slot count, workspace size, a pinned host buffer for the counts and the event that guards it.
    A call with known shapes then costs: patch the level pointers, allocate the two result tensors, ONE ctypes call,
    one pinned device->host copy of the counts, one `split`."""

    def __init__(self, levels, num_anchors, thr, nms_thres, decode, anchors, strides, layout, scale_x_y):
        lib = L.load()
        self.d = _yolo_desc(levels, num_anchors, thr, nms_thres, decode, anchors, strides, layout, scale_x_y)
        n, n_pad = ctypes.c_int32(), ctypes.c_int32()
        L.check(lib.b200det_yolo_num_candidates(ctypes.byref(self.d), ctypes.byref(n), ctypes.byref(n_pad)), "yolo_num_candidates")
        self.n_pad, self.B = n_pad.value, self.d.batch
        self.ws_bytes = lib.b200det_yolo_workspace_bytes(ctypes.byref(self.d))
        self.dref = ctypes.byref(self.d)
        # count [B] | offsets [B+1], written by the device straight into this pinned (mapped, UVA) host buffer as soon as
        # the NMS stage is done; `event` is recorded by the library at that point, before the emit stage
        self.host = torch.empty((2 * self.B + 1,), dtype=torch.int32).pin_memory()
        self.event = torch.cuda.Event()
        self.event.record()                               # creates the underlying cudaEvent_t
        self.host_ptr = self.host.data_ptr()
        self.event_ptr = self.event.cuda_event
        self.finish = L.hostglue().finish_views
        self.fn = lib.b200det_yolo_nms_early
        self.lock = threading.Lock()                     # the descriptor and the pinned counts are per plan, not per call
        self.spare = None                                # result buffers allocated ahead for the next call (while the GPU was busy)


_yolo_plans = collections.OrderedDict()
_YOLO_PLANS_MAX = 16


def _plan_key(levels, num_anchors, thr, nms_thres, decode, anchors, strides, layout, scale_x_y):
    def freeze(x):
        if x is None:
            return None
        return tuple(tuple(float(v) for v in torch.as_tensor(a, dtype=torch.float32).reshape(-1).tolist()) if not isinstance(a, (int, float))
                     else float(a) for a in x)
    t0 = levels[0]
    return (t0.device.index, tuple(t.shape for t in levels), num_anchors, float(thr), float(nms_thres), decode,
            freeze(anchors), freeze(strides), layout, float(scale_x_y))


def _yolo_nms(predictions, num_anchors, conf_thres, nms_thres, compat, decode, anchors, strides, return_index, layout=None,
              scale_x_y=1.0):
    if not isinstance(predictions, (list, tuple)):
        predictions = [predictions]                      # model/YOLOV3.py:281-282
    thr = YOLO_FORCED_CONF_THRES if compat else conf_thres
    if len(predictions) == 0 or len(predictions) > L.MAX_LEVELS:
        raise ValueError(f"need 1..{L.MAX_LEVELS} prediction levels, got {len(predictions)}")
    for i, t in enumerate(predictions):
        L.require_cuda(t, f"predictions[{i}]")
    key = _plan_key(predictions, num_anchors, thr, nms_thres, decode, anchors, strides, layout, scale_x_y)
    plan = _yolo_plans.get(key)
    if plan is None:
        with torch.cuda.device(predictions[0].device):   # the plan's event belongs to the device of the heads
            plan = _yolo_plans[key] = _YoloPlan(predictions, num_anchors, thr, nms_thres, decode, anchors, strides, layout, scale_x_y)
        while len(_yolo_plans) > _YOLO_PLANS_MAX:
            _yolo_plans.popitem(last=False)
    d, B, n_pad = plan.d, plan.B, plan.n_pad
    dev = predictions[0].device
    with plan.lock:
        return _yolo_nms_planned(plan, predictions, d, B, n_pad, dev, return_index)


_trace = None        # tools/api_profile.py installs a list here: perf_counter stamps at the phase boundaries of the call


def _yolo_nms_planned(plan, predictions, d, B, n_pad, dev, return_index):
    tr = _trace
    if tr is not None:
        tr.append(time.perf_counter())
    for i, t in enumerate(predictions):                  # same shapes as the planned call (part of the key): only pointers move
        if t.device != dev:
            raise ValueError("all prediction levels must live on one device")
        if not t.is_contiguous():
            raise ValueError(f"predictions[{i}] must be contiguous (the reference .view()s it, model/YOLOV3.py:296)")
        d.head[i] = t.data_ptr()
    if torch.cuda.current_device() != dev.index:        # allocations below go to the current device: switch only when needed
        with torch.cuda.device(dev):
            return _yolo_nms_planned(plan, predictions, d, B, n_pad, dev, return_index)
    stream = torch.cuda.current_stream(dev).cuda_stream
    ws = L.workspace(plan.ws_bytes, dev, stream)
    spare, plan.spare = plan.spare, None
    if spare is not None and spare[2] == stream:         # allocated under this very stream (and, by the plan key, this device)
        rows, count = spare[0], spare[1]
    else:
        rows = torch.empty((B, n_pad, 7), dtype=torch.float32, device=dev)
        count = torch.empty((B,), dtype=torch.int32, device=dev)
    index = torch.empty((B, n_pad), dtype=torch.int32, device=dev) if return_index else None
    if tr is not None:
        tr.append(time.perf_counter())
    L.check(plan.fn(plan.dref, ws.data_ptr(), ws.numel(), rows.data_ptr(), index.data_ptr() if return_index else None,
                    count.data_ptr(), plan.host_ptr, plan.event_ptr, stream), "yolo_nms_early")
    if tr is not None:
        tr.append(time.perf_counter())
    # ---- the GPU is busy for the next few hundred microseconds: everything that does not need the counts happens now ----
    views = list(rows.unbind(0))                         # B views [n_pad, 7]; shrunk in place once the counts are known
    iviews = list(index.long().unbind(0)) if return_index else None
    plan.spare = (torch.empty((B, n_pad, 7), dtype=torch.float32, device=dev),           # the next call's result buffers
                  torch.empty((B,), dtype=torch.int32, device=dev), stream)
    if tr is not None:
        tr.append(time.perf_counter())
    # the one host sync of the call: it waits for the NMS stage only (the counts are final there and were written straight
    # into pinned host memory).  The emit kernel may still be writing the rows when this function returns; whatever the
    # caller does with them next is stream-ordered behind it, as with any torch op.
    plan.event.synchronize()
    if tr is not None:
        tr.append(time.perf_counter())
    # B shrinks in one call (csrc/hostglue.cpp): images without detections become None (YOLOV3.py:306,333)
    out: List[Optional[torch.Tensor]] = plan.finish(views, plan.host_ptr, 7)
    if tr is not None:
        tr.append(time.perf_counter())
    if return_index:
        return out, plan.finish(iviews, plan.host_ptr, 0)
    return out


def non_max_suppression(self, predictions, conf_thres=0.5, nms_thres=0.4, *, compat=True, decode=None, anchors=None,
                        strides=None, return_index=False, layout=None, scale_x_y=1.0):
    """Drop-in for YOLOv3/v4/v5 `non_max_suppression` (3 anchors per level).

    Returns a list (one entry per image) of `None` or fp32 `[K,7]` rows
    `(x1, y1, x2, y2, object_conf, class_score, class_pred)` in descending score order.
    compat=True (default) reproduces the reference bit-for-bit, including its forced
    `conf_thres = -0.0151`; compat=False honours `conf_thres`.  Extensions (keyword-only):
    `decode` in {None,'yolo_exp','yolov5','yolov4_norm'} with per-level `anchors`/`strides` (`yolov4_norm`: grid-unit
    anchors, `scale_x_y`, boxes come out as normalised corners — utils/YoloV4Utils.py:36-176), `return_index`, and
    `layout='channels_last'`: the levels are read as what their `[B, A, G, G, 5+C]` shape says (the layout YOLOv5's
    head really writes, model/YOLOV5.py:96) instead of the reference's planar re-interpretation of the same bytes
    (YOLOV5.py:178-183) — same result as the default on the permuted tensor, without the permute copy.
    """
    return _yolo_nms(predictions, 3, conf_thres, nms_thres, compat, decode, anchors, strides, return_index, layout, scale_x_y)


class _HostPipe:
    """Buffers and streams of the host-input pipeline for one (device, per-image head shapes, chunk) configuration."""

    def __init__(self, dev, shapes, chunk, n_pad):
        self.s_in, self.s_cmp = (torch.cuda.Stream(dev) for _ in range(2))
        self.dev_in = [[torch.empty((chunk,) + tuple(sh[1:]), dtype=torch.float32, device=dev) for sh in shapes]
                       for _ in range(2)]
        self.buf_free = [None, None]      # event: the pipeline has consumed what dev_in[i] last held
        self.chunk_seq = 0                # chunks submitted so far (across calls): dev_in alternates on it
        self.calls = 0                    # calls submitted so far: the two pinned result sets alternate on it
        B = shapes[0][0]
        self.batch, self.chunk, self.n_pad, self.last_done = B, chunk, n_pad, None
        nchunks = (B + chunk - 1) // chunk
        # Two result sets in PINNED (mapped) host memory.  The emit kernel writes the kept rows of a chunk straight into them
        # (packed, chunk c in the region starting at row c * chunk * n_pad) and the prefix kernel the chunk's count | offsets
        # words: no device->host copy is enqueued at all, and only kept rows cross the link.
        self.host = [(torch.empty((B * n_pad, 7), dtype=torch.float32).pin_memory(),
                      torch.empty((B * n_pad,), dtype=torch.int32).pin_memory(),
                      torch.zeros((nchunks, 2 * chunk + 1), dtype=torch.int32).pin_memory()) for _ in range(2)]
        self.dev_meta = [torch.empty((2 * chunk + 1,), dtype=torch.int32, device=dev) for _ in range(2)]

    def in_flight(self):
        return self.last_done is not None and not self.last_done.query()


_host_pipes = collections.OrderedDict()      # LRU over (device, per-image head shapes, chunk, anchors)
_HOST_PIPES_MAX = 4


def clear_host_pipes():
    """Release the pinned result buffers, device input buffers and streams of the host-input pipelines."""
    _host_pipes.clear()


class HostNmsHandle:
    """A submitted `non_max_suppression_host_async` call; `result()` waits for the last chunk's pipeline."""

    def __init__(self, done, host, return_index, chunks, pipe, generation, batch):
        self._done, self._host, self._return_index, self._chunks = done, host, return_index, chunks
        self._pipe, self._generation, self._batch = pipe, generation, batch
        self._out = None

    def result(self):
        if self._out is None:
            # the rows live in one of the pipe's two pinned sets: a third submission before this call was collected has
            # re-used the set — fail loudly instead of returning another batch's rows
            if self._pipe.calls - self._generation > 2:
                raise RuntimeError("non_max_suppression_host_async: this call's pinned result buffer was overwritten by a "
                                   "later submission (at most two calls may be in flight per configuration)")
            self._done.synchronize()
            rows, index, meta = self._host
            n_pad = self._pipe.n_pad
            out: List[Optional[torch.Tensor]] = []
            idx: List[Optional[torch.Tensor]] = []
            for c, (lo, hi) in enumerate(self._chunks):
                nb = hi - lo
                m = meta[c].tolist()                       # count [nb] | offsets [nb + 1], written by the device
                base = lo * n_pad
                for i in range(nb):
                    k, o = m[i], base + m[nb + i]
                    out.append(rows[o:o + k] if k else None)
                    if self._return_index:
                        idx.append(index[o:o + k].long() if k else None)
            self._out = (out, idx) if self._return_index else out
        return self._out


def non_max_suppression_host_async(self, predictions, conf_thres=0.5, nms_thres=0.4, *, compat=True, num_anchors=3,
                                   device=None, chunk_images=8, decode=None, anchors=None, strides=None, return_index=False):
    """Submit `non_max_suppression_host` without waiting: returns a `HostNmsHandle` whose `result()` is that function's
    return value.  Submitting batch i+1 before collecting batch i keeps the host->device link busy across batches (the
    tail of a batch — the pipeline of its last chunk — otherwise leaves it idle).  Two calls may be in flight per
    (device, shapes) configuration: the rows of a call live in one of two pinned buffers and stay valid until the second
    submission after it.  `predictions` must not be modified before `result()` returns."""
    if not isinstance(predictions, (list, tuple)):
        predictions = [predictions]
    for i, t in enumerate(predictions):
        if not isinstance(t, torch.Tensor) or t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise TypeError(f"predictions[{i}] must be a contiguous fp32 HOST tensor (use non_max_suppression for CUDA tensors)")
    if not torch.cuda.is_available():
        raise RuntimeError("b200det has no CPU path: non_max_suppression_host needs a CUDA device to run on")
    lib = L.load()
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    B = predictions[0].shape[0]
    chunk = max(1, min(int(chunk_images), B))
    thr = YOLO_FORCED_CONF_THRES if compat else conf_thres
    shapes = tuple(tuple(t.shape) for t in predictions)
    # keyed on the PER-IMAGE shapes: a smaller batch (the last one of an epoch) re-uses the buffers of the largest seen
    key = (dev.index, tuple(sh[1:] for sh in shapes), int(chunk_images), num_anchors)
    pipe = _host_pipes.get(key)
    if pipe is not None and pipe.batch < B:
        if pipe.in_flight():
            torch.cuda.synchronize(dev)
        del _host_pipes[key]
        pipe = None
    if pipe is None:
        # slots per image: every level padded to whole tiles (b200det_yolo_num_candidates)
        n_pad = sum((num_anchors * t.shape[2] * t.shape[2] + L.TILE - 1) // L.TILE * L.TILE for t in predictions)
        with torch.cuda.device(dev):
            pipe = _HostPipe(dev, shapes, max(1, min(int(chunk_images), B)), n_pad)
        _host_pipes[key] = pipe
        while len(_host_pipes) > _HOST_PIPES_MAX:
            _host_pipes.popitem(last=False)
    else:
        _host_pipes.move_to_end(key)
    chunk = min(chunk, pipe.chunk)
    nchunks = (B + chunk - 1) // chunk
    host = pipe.host[pipe.calls & 1]
    pipe.calls += 1
    generation = pipe.calls
    n_pad = pipe.n_pad
    rows_h, index_h, meta_h = host
    chunks = []
    with torch.cuda.device(dev):
        for c in range(nchunks):
            lo, hi = c * chunk, min(B, (c + 1) * chunk)
            nb = hi - lo
            chunks.append((lo, hi))
            slot = pipe.chunk_seq & 1
            pipe.chunk_seq += 1
            dst = [t[:nb] for t in pipe.dev_in[slot]]
            ev_in, ev_done = torch.cuda.Event(), torch.cuda.Event()
            with torch.cuda.stream(pipe.s_in):
                if pipe.buf_free[slot] is not None:
                    pipe.s_in.wait_event(pipe.buf_free[slot])        # the pipeline is done with this input buffer
                for d_, t in zip(dst, predictions):
                    d_.copy_(t[lo:hi], non_blocking=True)
                ev_in.record(pipe.s_in)
            with torch.cuda.stream(pipe.s_cmp):
                pipe.s_cmp.wait_event(ev_in)
                d = _yolo_desc(dst, num_anchors, thr, nms_thres, decode, anchors, strides)
                ws_bytes = lib.b200det_yolo_workspace_bytes(ctypes.byref(d))
                ws = L.workspace(ws_bytes, dev)
                dm = pipe.dev_meta[slot]
                L.check(lib.b200det_yolo_nms_packed(ctypes.byref(d), ws.data_ptr(), ws.numel(),
                                                    rows_h.data_ptr() + lo * n_pad * 28,
                                                    index_h.data_ptr() + lo * n_pad * 4 if return_index else None,
                                                    dm.data_ptr(), dm.data_ptr() + 4 * nb, meta_h[c].data_ptr(), None,
                                                    L.stream_ptr(dev)), "yolo_nms_packed")
                ev_done.record(pipe.s_cmp)
            pipe.buf_free[slot] = ev_done
        done = ev_done
    pipe.last_done = done
    return HostNmsHandle(done, host, return_index, chunks, pipe, generation, B)


def non_max_suppression_host(self, predictions, conf_thres=0.5, nms_thres=0.4, *, compat=True, num_anchors=3,
                             device=None, chunk_images=8, decode=None, anchors=None, strides=None, return_index=False):
    """`non_max_suppression` for predictions that live in HOST memory (pinned memory for full PCIe speed).

    Same arguments and the same rows as `non_max_suppression` (model/YOLOV5.py:157); the batch is cut into chunks of
    `chunk_images` images and the legs run on separate streams, so the host->device copy of chunk k+1 overlaps
    the CUDA pipeline of chunk k and the device->host copy of chunk k-1 (every image is independent).
    The detections are not copied back: the emit kernel writes the kept rows, packed, straight into pinned (mapped) host
    memory, so only kept rows cross the link and no device->host copy is enqueued.
    Returns a list of `None` / fp32 `[K,7]` HOST tensors: views into one of two pinned buffers that alternate between
    calls with the same shapes (clone them to keep them beyond the next call but one)."""
    return non_max_suppression_host_async(self, predictions, conf_thres, nms_thres, compat=compat, num_anchors=num_anchors,
                                          device=device, chunk_images=chunk_images, decode=decode, anchors=anchors,
                                          strides=strides, return_index=return_index).result()


def non_max_suppression_v2(self, predictions, conf_thres=0.5, nms_thres=0.4, *, compat=True, decode=None, anchors=None,
                           strides=None, return_index=False):
    """Drop-in for YOLOv2 `non_max_suppression` (5 anchors, model/YOLOV2.py:179-183)."""
    return _yolo_nms(predictions, 5, conf_thres, nms_thres, compat, decode, anchors, strides, return_index)


def prior_nms_raw(loc: torch.Tensor, cls: torch.Tensor, priors: torch.Tensor, topk=100, nms_thresh=0.5,
                  class_thresh=0.45, mode="union", compat=True, want_index=False, count_out: Optional[torch.Tensor] = None):
    """Enqueue the prior pipeline; returns device tensors `(rows [B, topk, 7], index [B, topk] | None, count [2, B])` without
    synchronising.  `count_out`: an int32 `[2, B]` buffer to receive the counts instead of a fresh device tensor — it may be
    pinned (mapped) HOST memory, which the kernels then write directly."""
    if mode not in ("union", "min"):
        raise TypeError("Unknown nms mode: %s." % mode)  # model/SSD.py:298-299
    lib = L.load()
    L.require_cuda(loc, "loc_preds"); L.require_cuda(cls, "cls_preds"); L.require_cuda(priors, "iou_boxes")
    if loc.dim() != 3 or loc.shape[2] != 4 or cls.dim() != 3 or cls.shape[:2] != loc.shape[:2] or \
            tuple(priors.shape) != (loc.shape[1], 4):
        raise ValueError(f"expected loc [B,P,4], cls [B,P,C], priors [P,4]; got {tuple(loc.shape)}, {tuple(cls.shape)}, "
                         f"{tuple(priors.shape)}")
    loc, cls, priors = loc.contiguous(), cls.contiguous(), priors.contiguous()
    dev = loc.device
    d = L.PriorDesc()
    d.batch, d.num_priors, d.num_classes = loc.shape[0], loc.shape[1], cls.shape[2]
    d.loc, d.cls, d.priors = loc.data_ptr(), cls.data_ptr(), priors.data_ptr()
    d.topk, d.nms_thresh, d.class_thresh = int(topk), float(nms_thresh), float(class_thresh)
    d.mode_min, d.compat = int(mode == "min"), int(bool(compat))
    ws_bytes = lib.b200det_prior_workspace_bytes(ctypes.byref(d))
    with torch.cuda.device(dev):
        ws = L.workspace(ws_bytes, dev)
        rows = torch.empty((d.batch, d.topk, 7), dtype=torch.float32, device=dev)
        index = torch.empty((d.batch, d.topk), dtype=torch.int32, device=dev) if want_index else None
        count = count_out if count_out is not None else torch.empty((2, d.batch), dtype=torch.int32, device=dev)
        L.check(lib.b200det_prior_nms(ctypes.byref(d), ws.data_ptr(), ws.numel(), rows.data_ptr(),
                                      index.data_ptr() if want_index else None, count[0].data_ptr(), count[1].data_ptr(),
                                      L.stream_ptr(dev)), "prior_nms")
    return rows, index, count


class _PriorPlan:
    """What a `prior_non_max_suppression` call needs beyond the three pointers, per (device, shapes, options): the descriptor,
    the workspace size, a pinned host buffer the kernels write the two count rows into, the event behind them."""

    def __init__(self, loc, cls, topk, nms_thresh, class_thresh, mode, compat):
        lib = L.load()
        d = self.d = L.PriorDesc()
        d.batch, d.num_priors, d.num_classes = loc.shape[0], loc.shape[1], cls.shape[2]
        d.topk, d.nms_thresh, d.class_thresh = int(topk), float(nms_thresh), float(class_thresh)
        d.mode_min, d.compat = int(mode == "min"), int(bool(compat))
        self.dref = ctypes.byref(d)
        self.ws_bytes = lib.b200det_prior_workspace_bytes(self.dref)
        self.B, self.topk = d.batch, d.topk
        self.host = torch.empty((2, self.B), dtype=torch.int32).pin_memory()      # kept rows [0] | candidates above the threshold [1]
        self.host_ptr = self.host.data_ptr()
        self.event = torch.cuda.Event()
        self.fn = lib.b200det_prior_nms
        self.finish = L.hostglue().finish_views
        self.lock = threading.Lock()


_prior_plans = collections.OrderedDict()


def prior_non_max_suppression(self, predictions, topk=100, nms_thresh=0.5, class_thresh=0.45, mode="union", *,
                              compat=True, return_index=False):
    """Drop-in for SSD / RetinaNet `non_max_suppression` (model/SSD.py:249, model/RetinaNet.py:117).
    Priors come from `self.iou_boxes` [P,4].  Returns a list of fp32 `[K,7]` rows
    `(x1, y1, x2, y2, 0, score, label)`; compat=True reproduces the reference's quirks (the last
    surviving box is dropped, boxes/labels gathered with score-filtered indices, IndexError when exactly
    one candidate passes the score threshold)."""
    if mode not in ("union", "min"):
        raise TypeError("Unknown nms mode: %s." % mode)  # model/SSD.py:298-299
    loc, cls = predictions
    priors = self.iou_boxes
    L.require_cuda(loc, "loc_preds"); L.require_cuda(cls, "cls_preds"); L.require_cuda(priors, "iou_boxes")
    if loc.dim() != 3 or loc.shape[2] != 4 or cls.dim() != 3 or cls.shape[:2] != loc.shape[:2] or \
            tuple(priors.shape) != (loc.shape[1], 4):
        raise ValueError(f"expected loc [B,P,4], cls [B,P,C], priors [P,4]; got {tuple(loc.shape)}, {tuple(cls.shape)}, "
                         f"{tuple(priors.shape)}")
    dev = loc.device
    key = (dev.index, loc.shape, cls.shape, int(topk), float(nms_thresh), float(class_thresh), mode, bool(compat))
    plan = _prior_plans.get(key)
    if plan is None:
        with torch.cuda.device(dev):
            plan = _prior_plans[key] = _PriorPlan(loc, cls, topk, nms_thresh, class_thresh, mode, compat)
        while len(_prior_plans) > _YOLO_PLANS_MAX:
            _prior_plans.popitem(last=False)
    if torch.cuda.current_device() != dev.index:
        with torch.cuda.device(dev):
            return prior_non_max_suppression(self, predictions, topk, nms_thresh, class_thresh, mode, compat=compat,
                                             return_index=return_index)
    loc, cls, priors = loc.contiguous(), cls.contiguous(), priors.contiguous()
    B, d = plan.B, plan.d
    with plan.lock:                                      # the descriptor and the pinned counts are per plan, not per call
        d.loc, d.cls, d.priors = loc.data_ptr(), cls.data_ptr(), priors.data_ptr()
        stream = torch.cuda.current_stream(dev).cuda_stream
        ws = L.workspace(plan.ws_bytes, dev, stream)
        rows = torch.empty((B, plan.topk, 7), dtype=torch.float32, device=dev)
        index = torch.empty((B, plan.topk), dtype=torch.int32, device=dev) if return_index else None
        # kept rows [0] and candidates above the score threshold [1] are written by the device straight into pinned host memory
        L.check(plan.fn(plan.dref, ws.data_ptr(), ws.numel(), rows.data_ptr(), index.data_ptr() if return_index else None,
                        plan.host_ptr, plan.host_ptr + 4 * B, stream), "prior_nms")
        plan.event.record()
        views = list(rows.unbind(0))                     # built while the GPU works; shrunk in place after the one sync
        iviews = list(index.long().unbind(0)) if return_index else None
        plan.event.synchronize()
        if compat and 1 in plan.host[1].tolist():
            raise IndexError("too many indices for tensor of dimension 1")   # model/SSD.py:262,266 (0-dim index)
        out = plan.finish(views, plan.host_ptr, 7, True)
        if return_index:
            return out, plan.finish(iviews, plan.host_ptr, 0, True)
    return out


_prior_host_bufs = {}


def prior_non_max_suppression_host(self, predictions, topk=100, nms_thresh=0.5, class_thresh=0.45, mode="union", *,
                                   compat=True, device=None):
    """`prior_non_max_suppression` for `(loc, cls)` that live in HOST memory (pinned for full link speed): the two tensors
    are copied into device buffers kept per shape, the pipeline runs, and the `[B, topk, 7]` rows plus the counts come
    back into pinned host memory.  Returns a list of fp32 `[K,7]` HOST tensors (views of a pinned buffer that the next
    call with the same shapes re-uses).  The copy of the logits IS the step (1.29 GB for RetinaNet-800 at batch 32
    against 0.35 ms of kernels), so the copies are not chunked."""
    loc, cls = predictions
    for nm, t in (("loc_preds", loc), ("cls_preds", cls)):
        if not isinstance(t, torch.Tensor) or t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise TypeError(f"{nm} must be a contiguous fp32 HOST tensor (use prior_non_max_suppression for CUDA tensors)")
    if not torch.cuda.is_available():
        raise RuntimeError("b200det has no CPU path: prior_non_max_suppression_host needs a CUDA device to run on")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    key = (dev.index, tuple(loc.shape), tuple(cls.shape), int(topk))
    bufs = _prior_host_bufs.get(key)
    if bufs is None:
        if len(_prior_host_bufs) >= 4:
            _prior_host_bufs.clear()
        B = loc.shape[0]
        bufs = _prior_host_bufs[key] = (torch.empty(loc.shape, dtype=torch.float32, device=dev),
                                        torch.empty(cls.shape, dtype=torch.float32, device=dev),
                                        torch.empty((B, int(topk), 7), dtype=torch.float32).pin_memory(),
                                        torch.empty((2, B), dtype=torch.int32).pin_memory())
    dloc, dcls, hrows, hcount = bufs
    pri = self.iou_boxes if self.iou_boxes.is_cuda else self.iou_boxes.to(dev)
    with torch.cuda.device(dev):
        dloc.copy_(loc, non_blocking=True)
        dcls.copy_(cls, non_blocking=True)
        rows, _, count = prior_nms_raw(dloc, dcls, pri, topk, nms_thresh, class_thresh, mode, compat, False)
        hrows.copy_(rows, non_blocking=True)
        hcount.copy_(count, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
    if compat and bool((hcount[1] == 1).any()):
        raise IndexError("too many indices for tensor of dimension 1")   # model/SSD.py:262,266 (0-dim index)
    return [r[:k] for r, k in zip(hrows.unbind(0), hcount[0].tolist())]


def decode_box(head: torch.Tensor, anchors, stride: float, mode: str = "yolo_exp", num_anchors: Optional[int] = None,
               scale_x_y: float = 1.0):
    """Full decoded map of one level: planar `[B, A*(5+C), G, G]` -> `[B, A*G*G, 5+C]`.

    mode 'yolo_exp' (D1, accuracy.py:412-435,459-466): x=(σ+gx)·stride, w=exp·anchor·stride, σ(conf), σ(cls);
         `anchors` are the SCALED anchors (grid units) exactly as the reference's callers pass them.
    mode 'yolov5'  (D2, utils/YoloV5Utils.py:244-248): xy=(2σ-0.5+g)·stride, wh=(2σ)²·anchor (pixel anchors).
    mode 'none': the planar->rows permute of model/YOLOV3.py:294-300 only.
    mode 'yolov4_norm' (D3, utils/YoloV4Utils.py:36-176): rows `(x1, y1, x2, y2, σ(obj), σ(cls)·σ(obj) ...)` with
         normalised corner boxes; `anchors` in grid units, `scale_x_y` as there, `stride` unused
         (`yolo_forward_dynamic` is the same kernel with the reference's two-tensor return).
    """
    lib = L.load()
    L.require_cuda(head, "head")
    if not head.is_contiguous():
        raise ValueError("head must be contiguous")
    m = _DECODE[mode]
    anc = None
    if anchors is not None:
        anc = torch.as_tensor(anchors, dtype=torch.float32).reshape(-1, 2).to(head.device).contiguous()
        A = anc.shape[0]
    else:
        A = num_anchors
    if A is None:
        raise ValueError("need `anchors` or `num_anchors`")
    if m != L.DECODE_NONE and anc is None:
        raise ValueError("decode modes need anchors")
    B, G = head.shape[0], head.shape[2]
    if head.numel() % (B * A * G * G):
        raise ValueError(f"head of shape {tuple(head.shape)} is not [B, {A}, 5+C, {G}, {G}] storage")
    F = head.numel() // (B * A * G * G)
    out = torch.empty((B, A * G * G, F), dtype=torch.float32, device=head.device)
    if m == L.DECODE_YOLOV4_NORM:
        with torch.cuda.device(head.device):
            L.check(lib.b200det_yolo_forward_dynamic(head.data_ptr(), B, A, F - 5, G, G, anc.data_ptr(), float(scale_x_y),
                                                     out.data_ptr(), F, out.data_ptr() + 20, F, out.data_ptr() + 16, F,
                                                     L.stream_ptr(head.device)), "yolo_forward_dynamic")
        return out
    with torch.cuda.device(head.device):
        L.check(lib.b200det_decode_box(head.data_ptr(), B, A, F - 5, G, m, anc.data_ptr() if anc is not None else None,
                                       float(stride), out.data_ptr(), L.stream_ptr(head.device)), "decode_box")
    return out


def yolo_forward_dynamic(output, conf_thresh, num_classes, anchors, num_anchors, scale_x_y, only_objectness=1,
                         validation=False, *, return_det_confs=False):
    """Drop-in for `yolo_forward_dynamic` (LightningFunc/utils/YoloV4Utils.py:36-176), decode D3 of one YOLOv4 head level.

    `output` planar `[B, num_anchors*(5+num_classes), H, W]`; `anchors` the flat list `[w0, h0, w1, h1, ...]` in grid
    units the reference's callers pass (or an `[A,2]` tensor).  Returns `(boxes [B, A*H*W, 1, 4], confs [B, A*H*W, C])`:
    normalised corner boxes and `sigmoid(cls) * sigmoid(obj)`.  `conf_thresh`, `only_objectness` and `validation` are
    accepted and unused, as in the reference."""
    lib = L.load()
    L.require_cuda(output, "output")
    if output.dim() != 4 or output.shape[1] != num_anchors * (5 + num_classes):
        raise ValueError(f"output must be [B, {num_anchors}*(5+{num_classes}), H, W], got {tuple(output.shape)}")
    head = output.contiguous()
    B, H, W = head.shape[0], head.shape[2], head.shape[3]
    anc = torch.as_tensor(anchors, dtype=torch.float32).reshape(-1)[:2 * num_anchors].reshape(num_anchors, 2)
    anc = anc.to(head.device).contiguous()
    N = num_anchors * H * W
    boxes = torch.empty((B, N, 1, 4), dtype=torch.float32, device=head.device)
    confs = torch.empty((B, N, num_classes), dtype=torch.float32, device=head.device)
    det = torch.empty((B, N), dtype=torch.float32, device=head.device) if return_det_confs else None
    with torch.cuda.device(head.device):
        L.check(lib.b200det_yolo_forward_dynamic(head.data_ptr(), B, num_anchors, num_classes, H, W, anc.data_ptr(),
                                                 float(scale_x_y), boxes.data_ptr(), 4, confs.data_ptr(), num_classes,
                                                 det.data_ptr() if det is not None else None, 1, L.stream_ptr(head.device)),
                "yolo_forward_dynamic")
    return (boxes, confs, det) if return_det_confs else (boxes, confs)


def get_region_boxes(boxes_and_confs):
    """Drop-in for `get_region_boxes` (LightningFunc/utils/YoloV4Utils.py:18-34): concatenates the per-level results of
    `yolo_forward_dynamic` along the candidate axis."""
    return [torch.cat([item[0] for item in boxes_and_confs], dim=1), torch.cat([item[1] for item in boxes_and_confs], dim=1)]

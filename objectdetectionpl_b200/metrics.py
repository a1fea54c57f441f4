"""Detection metrics with the reference's call signatures, computed on the GPU:

  get_batch_statistics(outputs, targets, iou_threshold)      replaces LightningFunc/accuracy.py:116-154
  ap_per_class(tp, conf, pred_cls, target_cls)               replaces LightningFunc/accuracy.py:207-260 (+ compute_ap :262)

  get_yolo_statistics(self, output, target)                  replaces LightningFunc/accuracy.py:382-470 (YOLOv2..v4 test step)

They return numpy objects of the reference's shapes and dtypes, so `LightningFunc/step.py:95,99,115` runs unchanged;
`batch_statistics_raw` / `ap_per_class_device` are the device-resident building blocks (no host sync).
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L


def batch_statistics_raw(rows: torch.Tensor, row_start: torch.Tensor, count: torch.Tensor, max_count: int,
                         targets: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """`tp` (fp32, one slot per row of `rows`, 1.0 = true positive) for detections laid out as
    rows[row_start[b] : row_start[b] + count[b]] per image b.  Everything stays on the device."""
    lib = L.load()
    L.require_cuda(rows, "rows")
    L.require_cuda(targets, "targets")
    L.require_cuda(row_start, "row_start", torch.int64)
    L.require_cuda(count, "count", torch.int32)
    if rows.dim() != 2 and rows.dim() != 3 or rows.shape[-1] != 7 or not rows.is_contiguous():
        raise ValueError(f"rows must be contiguous [..., 7]; got {tuple(rows.shape)}")
    if targets.dim() != 2 or targets.shape[1] != 6:
        raise ValueError(f"targets must be [nt, 6] (image, label, x1, y1, x2, y2); got {tuple(targets.shape)}")
    targets = targets.contiguous()
    B, nt = int(count.shape[0]), int(targets.shape[0])
    dev = rows.device
    tp = torch.zeros(rows.shape[:-1], dtype=torch.float32, device=dev)
    nb = lib.b200det_batch_statistics_workspace_bytes(B, nt)
    with torch.cuda.device(dev):
        ws = L.workspace(nb, dev)
        L.check(lib.b200det_batch_statistics(rows.data_ptr(), row_start.data_ptr(), count.data_ptr(), B, int(max_count),
                                             targets.data_ptr() if nt else None, nt, float(iou_threshold), ws.data_ptr(),
                                             ws.numel(), tp.data_ptr(), L.stream_ptr(dev)), "batch_statistics")
    return tp


def get_batch_statistics(outputs: Sequence[Optional[torch.Tensor]], targets: torch.Tensor, iou_threshold: float):
    """Drop-in for `get_batch_statistics` (accuracy.py:116): per image with detections a list
    `[true_positives float64 [K], pred_scores [K], pred_labels [K]]` of numpy arrays (`None` outputs are skipped)."""
    live = [(i, o) for i, o in enumerate(outputs) if o is not None]
    if not live:
        return []
    dev = live[0][1].device
    L.require_cuda(live[0][1], "outputs")
    B = len(outputs)
    counts = [0 if o is None else int(o.shape[0]) for o in outputs]
    packed = torch.cat([o for _, o in live], 0).contiguous()
    starts = np.concatenate(([0], np.cumsum(counts)[:-1])).astype(np.int64)
    row_start = torch.from_numpy(starts).to(dev)
    count = torch.tensor(counts, dtype=torch.int32, device=dev)
    tp = batch_statistics_raw(packed, row_start, count, max(counts), targets.to(dev, torch.float32), iou_threshold)
    tp_h = tp.double().cpu().numpy()
    sc_h = packed[:, 4].cpu().numpy()
    lb_h = packed[:, 6].cpu().numpy()
    out = []
    for i, _ in live:
        s, e = int(starts[i]), int(starts[i]) + counts[i]
        out.append([tp_h[s:e], sc_h[s:e], lb_h[s:e]])
    return out


def ap_per_class_device(tp: torch.Tensor, conf: torch.Tensor, pred_cls: torch.Tensor, classes: torch.Tensor,
                        n_gt: torch.Tensor):
    """Device-resident core of `ap_per_class`: returns fp64 tensors (p, r, ap, f1), one entry per `classes`."""
    lib = L.load()
    for t, nm in ((tp, "tp"), (conf, "conf"), (pred_cls, "pred_cls")):
        L.require_cuda(t, nm)
    L.require_cuda(classes, "classes", torch.int32)
    L.require_cuda(n_gt, "n_gt", torch.int32)
    n, U = int(tp.shape[0]), int(classes.shape[0])
    if n > (1 << 30):
        raise ValueError(f"ap_per_class: {n} detections > 2^30")
    dev = tp.device
    out = torch.zeros((4, max(U, 1)), dtype=torch.float64, device=dev)
    nb = lib.b200det_ap_per_class_workspace_bytes(n)
    with torch.cuda.device(dev):
        ws = L.workspace(nb, dev)
        L.check(lib.b200det_ap_per_class(tp.contiguous().data_ptr(), conf.contiguous().data_ptr(),
                                         pred_cls.contiguous().data_ptr(), n, classes.data_ptr(), n_gt.data_ptr(), U,
                                         ws.data_ptr(), ws.numel(), out[0].data_ptr(), out[1].data_ptr(),
                                         out[2].data_ptr(), out[3].data_ptr(), L.stream_ptr(dev)), "ap_per_class")
    return out[0, :U], out[1, :U], out[2, :U], out[3, :U]


def ap_per_class(tp, conf, pred_cls, target_cls, device=None):
    """Drop-in for `ap_per_class` (accuracy.py:207): numpy in (as `test_epoch_end` passes them, step.py:112-115),
    numpy out `(p, r, ap, f1, unique_classes int32)`.  CUDA tensors are accepted as well."""
    if not torch.cuda.is_available():
        raise RuntimeError("b200det has no CPU path: ap_per_class needs a CUDA device")
    dev = torch.device(device) if device is not None else (
        tp.device if isinstance(tp, torch.Tensor) and tp.is_cuda else torch.device("cuda", torch.cuda.current_device()))

    def dev32(x):
        if isinstance(x, torch.Tensor):
            return x.to(dev, torch.float32)
        return torch.as_tensor(np.asarray(x, dtype=np.float32), device=dev)

    t_cls = np.asarray(target_cls.cpu() if isinstance(target_cls, torch.Tensor) else target_cls)
    classes, n_gt = np.unique(t_cls, return_counts=True)                 # accuracy.py:225, :231
    ok = (classes >= 0) & (classes < (1 << 30)) & (classes == np.floor(classes))
    if not ok.all():
        raise ValueError("ap_per_class: class ids must be integers in [0, 2^30)")
    p, r, ap, f1 = ap_per_class_device(dev32(tp), dev32(conf), dev32(pred_cls),
                                       torch.as_tensor(classes.astype(np.int32), device=dev),
                                       torch.as_tensor(n_gt.astype(np.int32), device=dev))
    res = torch.stack([p, r, ap, f1]).cpu().numpy()
    return res[0], res[1], res[2], res[3], classes.astype("int32")


def yolo_statistics_level(head: torch.Tensor, scaled_anchors: torch.Tensor, stride: float, target: torch.Tensor,
                          ignore_thres: float):
    """One level of `get_yolo_statistics` on the device: returns `(metrics fp32 [6], output [B, A*G*G, 5+C])` where
    metrics = (cls_acc, recall50, recall75, precision, conf_obj, conf_noobj) and `output` is the decoded map
    (boxes in pixels).  No host sync."""
    lib = L.load()
    L.require_cuda(head, "output")
    if not head.is_contiguous():
        raise ValueError("output must be contiguous (the reference .view()s it, accuracy.py:406)")
    dev = head.device
    an = scaled_anchors.to(dev, torch.float32).contiguous()
    tg = target.to(dev, torch.float32).contiguous()
    A, B, G = int(an.shape[0]), int(head.shape[0]), int(head.shape[2])
    if head.numel() % (B * A * G * G):
        raise ValueError(f"output of shape {tuple(head.shape)} is not [B, {A}*(5+C), {G}, {G}] storage")
    F = head.numel() // (B * A * G * G)
    nt = int(tg.shape[0])
    rows = torch.empty((B, A * G * G, F), dtype=torch.float32, device=dev)
    metrics = torch.empty((6,), dtype=torch.float32, device=dev)
    nb = lib.b200det_yolo_statistics_workspace_bytes(B, A, G, nt)
    with torch.cuda.device(dev):
        ws = L.workspace(nb, dev)
        L.check(lib.b200det_yolo_statistics_level(head.data_ptr(), B, A, F - 5, G, an.data_ptr(), float(stride),
                                                  tg.data_ptr() if nt else None, nt, float(ignore_thres), ws.data_ptr(),
                                                  ws.numel(), rows.data_ptr(), metrics.data_ptr(), L.stream_ptr(dev)),
                "yolo_statistics_level")
    return metrics, rows


def get_yolo_statistics(self, output, target):
    """Drop-in for `get_yolo_statistics` (accuracy.py:382): `{grid_size: [cls_acc, recall50, recall75, precision,
    conf_obj, conf_noobj (0-dim numpy), output (CPU tensor [B, A*G*G, 5+C])]}` per level.  Reads `self.anchors`,
    `self.anch_masks`, `self.num_classes`, `self.img_size`, `self.ignore_thres` and leaves the attributes the reference
    sets behind (`num_anchors`, `grid_size`, `stride`, `scaled_anchors`)."""
    batch_metrics = {}
    if type(output) != list:
        output = [output]
    for i, x in enumerate(output):
        if self.anch_masks is not None:                                   # YOLOv4 (:388-389)
            anchors = [self.anchors[m] for m in self.anch_masks[i]]
        elif len(self.anchors) == 3:                                      # YOLOv3 (:394-395)
            anchors = self.anchors[i]
        else:                                                             # YOLOv2 (:396-397)
            anchors = self.anchors
        self.num_anchors = len(anchors)
        g = x.size(2)
        self.grid_size = g
        self.stride = self.img_size / g                                   # :422
        scaled = torch.tensor([(a_w / self.stride, a_h / self.stride) for a_w, a_h in anchors], dtype=torch.float32,
                              device=x.device)                            # :427 (python-float division, then fp32)
        self.scaled_anchors = scaled
        if x.numel() // (x.size(0) * self.num_anchors * g * g) != self.num_classes + 5:
            raise RuntimeError(f"shape '{[x.size(0), self.num_anchors, self.num_classes + 5, g, g]}' is invalid for input "
                               f"of size {x.numel()}")                    # the reference's .view() error (:405-406)
        m, rows = yolo_statistics_level(x, scaled, self.stride, target, self.ignore_thres)
        mh = m.cpu().numpy()
        batch_metrics[g] = [mh[0], mh[1], mh[2], mh[3], mh[4], mh[5], rows.cpu()]
        batch_metrics[g][:6] = [np.asarray(v) for v in batch_metrics[g][:6]]
    return batch_metrics

"""TEST INFRASTRUCTURE ONLY — CPU oracle for the detection post-processing / target-assignment path.

This module is a *restatement* (torch-CPU eager ops, fp32) of the algorithms in
Leyan529/ObjectDetectionPL that the CUDA library replaces.  Every function cites the reference
`file:line` it follows.  It exists to CHECK the CUDA path; it must never be imported by the product
package `objectdetectionpl_b200/` (only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may use it).

Parity status: PINNED.  The reference ships no tests or golden vectors, so the pin is
made against the reference itself: `oracle/gen_golden.py` imports the unmodified reference in the
build container (through `oracle/ref_harness.py`), runs it on seeded inputs and stores inputs+outputs
under `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every function here against those
files, and `tests/test_oracle_vs_reference.py` re-checks live whenever `/root/reference` is present.

The arithmetic is torch-CPU because the reference's arithmetic IS torch eager ops (sigmoid/exp/atan,
`max`, `argsort`, masked gathers, `sum`); using the same primitives keeps last-ulp behaviour identical
(SURVEY.md §8c "third-party arithmetic").
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch

F32 = torch.float32

# --------------------------------------------------------------------------------------------
# N3 / N4 — box helpers
# --------------------------------------------------------------------------------------------

def xywh2xyxy(x: torch.Tensor) -> torch.Tensor:
    """Centre/size -> corner boxes.  Follows LightningFunc/accuracy.py:289-295."""
    half_w = x[..., 2] / 2
    half_h = x[..., 3] / 2
    return torch.stack((x[..., 0] - half_w, x[..., 1] - half_h, x[..., 0] + half_w, x[..., 1] + half_h), dim=-1)


def bbox_iou_plus1(box1: torch.Tensor, box2: torch.Tensor, x1y1x2y2: bool = True) -> torch.Tensor:
    """IoU with the '+1 pixel' convention and +1e-16 in the union.  LightningFunc/accuracy.py:39-69.

    box1 [1|n,4], box2 [n,4] -> [n].  The GIoU/DIoU/CIoU kwargs of the reference are accepted and
    ignored there (accuracy.py:39), so they do not exist here.
    """
    if not x1y1x2y2:  # accuracy.py:43-48
        a_x1, a_x2 = box1[:, 0] - box1[:, 2] / 2, box1[:, 0] + box1[:, 2] / 2
        a_y1, a_y2 = box1[:, 1] - box1[:, 3] / 2, box1[:, 1] + box1[:, 3] / 2
        b_x1, b_x2 = box2[:, 0] - box2[:, 2] / 2, box2[:, 0] + box2[:, 2] / 2
        b_y1, b_y2 = box2[:, 1] - box2[:, 3] / 2, box2[:, 1] + box2[:, 3] / 2
    else:  # accuracy.py:49-52
        a_x1, a_y1, a_x2, a_y2 = box1[:, 0], box1[:, 1], box1[:, 2], box1[:, 3]
        b_x1, b_y1, b_x2, b_y2 = box2[:, 0], box2[:, 1], box2[:, 2], box2[:, 3]
    ix1 = torch.max(a_x1, b_x1)  # accuracy.py:55-58
    iy1 = torch.max(a_y1, b_y1)
    ix2 = torch.min(a_x2, b_x2)
    iy2 = torch.min(a_y2, b_y2)
    inter = torch.clamp(ix2 - ix1 + 1, min=0) * torch.clamp(iy2 - iy1 + 1, min=0)  # accuracy.py:60-62
    area_a = (a_x2 - a_x1 + 1) * (a_y2 - a_y1 + 1)  # accuracy.py:64-65
    area_b = (b_x2 - b_x1 + 1) * (b_y2 - b_y1 + 1)
    return inter / (area_a + area_b - inter + 1e-16)  # accuracy.py:66-68


def bbox_iou_v5(box1: torch.Tensor, box2: torch.Tensor, x1y1x2y2: bool = True,
                GIoU: bool = False, DIoU: bool = False, CIoU: bool = False) -> torch.Tensor:
    """IoU / GIoU / DIoU / CIoU on transposed [4,n] boxes (no +1).  LightningFunc/accuracy.py:71-114."""
    if x1y1x2y2:  # accuracy.py:76-78
        a_x1, a_y1, a_x2, a_y2 = box1[0], box1[1], box1[2], box1[3]
        b_x1, b_y1, b_x2, b_y2 = box2[0], box2[1], box2[2], box2[3]
    else:  # accuracy.py:79-83
        a_x1, a_x2 = box1[0] - box1[2] / 2, box1[0] + box1[2] / 2
        a_y1, a_y2 = box1[1] - box1[3] / 2, box1[1] + box1[3] / 2
        b_x1, b_x2 = box2[0] - box2[2] / 2, box2[0] + box2[2] / 2
        b_y1, b_y2 = box2[1] - box2[3] / 2, box2[1] + box2[3] / 2
    inter = (torch.min(a_x2, b_x2) - torch.max(a_x1, b_x1)).clamp(0) * \
            (torch.min(a_y2, b_y2) - torch.max(a_y1, b_y1)).clamp(0)  # accuracy.py:86-87
    w1, h1 = a_x2 - a_x1, a_y2 - a_y1  # accuracy.py:90-92
    w2, h2 = b_x2 - b_x1, b_y2 - b_y1
    union = (w1 * h1 + 1e-16) + w2 * h2 - inter
    iou = inter / union  # accuracy.py:94
    if GIoU or DIoU or CIoU:
        cw = torch.max(a_x2, b_x2) - torch.min(a_x1, b_x1)  # accuracy.py:96-97
        ch = torch.max(a_y2, b_y2) - torch.min(a_y1, b_y1)
        if GIoU:  # accuracy.py:98-100
            c_area = cw * ch + 1e-16
            return iou - (c_area - union) / c_area
        c2 = cw ** 2 + ch ** 2 + 1e-16  # accuracy.py:103
        rho2 = ((b_x1 + b_x2) - (a_x1 + a_x2)) ** 2 / 4 + ((b_y1 + b_y2) - (a_y1 + a_y2)) ** 2 / 4  # :105
        if DIoU:
            return iou - rho2 / c2  # accuracy.py:107
        v = (4 / math.pi ** 2) * torch.pow(torch.atan(w2 / h2) - torch.atan(w1 / h1), 2)  # accuracy.py:109
        with torch.no_grad():
            alpha = v / (1 - iou + v)  # accuracy.py:110-111
        return iou - (rho2 / c2 + v * alpha)  # accuracy.py:112
    return iou


def pair_iou(t1: torch.Tensor, t2: torch.Tensor) -> torch.Tensor:
    """Elementwise corner-box IoU (no +1, no eps, areas clamped >= 0).  LightningFunc/accuracy.py:6-37."""
    lo = torch.max(t1[..., :2], t2[..., :2])  # accuracy.py:19-20
    hi = torch.min(t1[..., 2:], t2[..., 2:])
    d = torch.clamp(hi - lo, min=0.0)  # accuracy.py:22
    inter = d[..., 0] * d[..., 1]
    d1 = torch.clamp(t1[..., 2:] - t1[..., :2], min=0.0)  # accuracy.py:26-30
    d2 = torch.clamp(t2[..., 2:] - t2[..., :2], min=0.0)
    return inter / (d1[..., 0] * d1[..., 1] + d2[..., 0] * d2[..., 1] - inter)  # accuracy.py:32


def bbox_wh_iou(wh1: torch.Tensor, wh2: torch.Tensor) -> torch.Tensor:
    """IoU of origin-anchored (w,h) pairs: anchor wh1[2] vs targets wh2[nt,2].  accuracy.py:297-303."""
    w2, h2 = wh2[:, 0], wh2[:, 1]
    inter = torch.min(wh1[0], w2) * torch.min(wh1[1], h2)
    return inter / ((wh1[0] * wh1[1] + 1e-16) + w2 * h2 - inter)


# --------------------------------------------------------------------------------------------
# N1 — YOLO v2..v5 test-time merge-NMS
# --------------------------------------------------------------------------------------------

YOLO_FORCED_CONF_THRES = -0.0151  # model/YOLOV5.py:164, YOLOV3.py:280, YOLOV4.py:228, YOLOV2.py:166


def yolo_rows_from_planar(levels: Sequence[torch.Tensor], num_anchors: int) -> torch.Tensor:
    """Planar head memory -> [B, sum(A*G*G), 5+C] rows.  model/YOLOV3.py:289-303 (v5: YOLOV5.py:173-186).

    Each level's storage is read as [B, A, 5+C, G, G] whatever its nominal shape, exactly as the
    reference's `.view(B, A, answers, G, G)` does.
    """
    rows = []
    for t in levels:
        b = t.shape[0]
        g = t.shape[2]
        fields = t.numel() // (b * num_anchors * g * g)
        rows.append(t.reshape(b, num_anchors, fields, g, g).permute(0, 1, 3, 4, 2).reshape(b, -1, fields))
    return torch.cat(rows, dim=1)


def yolo_nms_rows(rows: torch.Tensor, conf_thres: float = YOLO_FORCED_CONF_THRES, nms_thres: float = 0.4,
                  return_index: bool = False):
    """Per-image class-aware merge-NMS on [B,N,5+C] rows (xywh in cols 0..3).  model/YOLOV3.py:304-335.

    Structure follows the reference loop one-to-one (sort by conf*max(cls), repeatedly take the top row,
    mark same-class rows with IoU_+1 > nms_thres, replace the top row's box by the conf-weighted mean of
    the marked rows, drop them).  With `return_index` the original candidate index of every kept row
    is returned as well (test-only extension).  Ties in the score are ordered by ascending candidate
    index (stable sort) — the rule this build publishes; the reference's own order on ties is
    implementation-defined.
    """
    rows = rows.clone()
    rows[..., :4] = xywh2xyxy(rows[..., :4])  # YOLOV3.py:305
    out: List[Optional[torch.Tensor]] = [None] * rows.shape[0]
    out_idx: List[Optional[torch.Tensor]] = [None] * rows.shape[0]
    for bi in range(rows.shape[0]):
        img = rows[bi]
        alive = img[:, 4] >= conf_thres  # YOLOV3.py:310
        cand = torch.nonzero(alive).flatten()
        img = img[alive]
        if img.shape[0] == 0:  # YOLOV3.py:312-313
            continue
        score = img[:, 4] * img[:, 5:].max(1)[0]  # YOLOV3.py:315
        order = torch.argsort(-score, stable=True)  # YOLOV3.py:317 (stable: published tie rule)
        img, cand = img[order], cand[order]
        cls_conf, cls_id = img[:, 5:].max(1, keepdim=True)  # YOLOV3.py:318
        det = torch.cat((img[:, :5], cls_conf.float(), cls_id.float()), 1)  # YOLOV3.py:319
        kept, kept_idx = [], []
        while det.shape[0]:  # YOLOV3.py:322-331
            hit = bbox_iou_plus1(det[0, :4].unsqueeze(0), det[:, :4]) > nms_thres
            hit &= det[0, -1] == det[:, -1]
            wgt = det[hit, 4:5]
            det[0, :4] = (wgt * det[hit, :4]).sum(0) / wgt.sum()
            kept.append(det[0])
            kept_idx.append(cand[0])
            det, cand = det[~hit], cand[~hit]
        out[bi] = torch.stack(kept)  # YOLOV3.py:332-333
        out_idx[bi] = torch.stack(kept_idx)
    return (out, out_idx) if return_index else out


def yolo_nms(levels, num_anchors: int = 3, conf_thres: float = 0.5, nms_thres: float = 0.4,
             compat: bool = True, return_index: bool = False):
    """Reference-signature entry: `conf_thres` is overwritten with -0.0151 in compat mode (YOLOV5.py:164)."""
    if not isinstance(levels, (list, tuple)):
        levels = [levels]  # YOLOV3.py:281-282
    thr = YOLO_FORCED_CONF_THRES if compat else conf_thres
    return yolo_nms_rows(yolo_rows_from_planar(levels, num_anchors), thr, nms_thres, return_index)


def yolo_nms_fast(rows: torch.Tensor, conf_thres: float = YOLO_FORCED_CONF_THRES, nms_thres: float = 0.4):
    """Equivalent O(sum n_c * K_c) formulation used for the larger parity cases (SURVEY.md §8c, 2nd
    equivalence): greedy class-aware NMS on the ORIGINAL corner boxes decides the keep set; every
    suppressed row is attributed to the first keeper that removed it; the keeper's box becomes the
    conf-weighted mean of its cluster (model/YOLOV3.py:322-331).  Returns (rows_list, index_list).
    Same arithmetic as `yolo_nms_rows` for the IoU test (bit-exact keep set); the cluster sums are
    accumulated with the same torch ops on the same row order, so boxes agree to the last ulp too.
    """
    rows = rows.clone()
    rows[..., :4] = xywh2xyxy(rows[..., :4])
    out, out_idx = [None] * rows.shape[0], [None] * rows.shape[0]
    for bi in range(rows.shape[0]):
        img = rows[bi]
        alive = img[:, 4] >= conf_thres
        cand = torch.nonzero(alive).flatten()
        img = img[alive]
        if img.shape[0] == 0:
            continue
        cls_conf, cls_id = img[:, 5:].max(1)
        score = img[:, 4] * cls_conf
        order = torch.argsort(-score, stable=True)
        img, cand, cls_conf, cls_id = img[order], cand[order], cls_conf[order], cls_id[order]
        n = img.shape[0]
        keep_rank, merged = [], {}
        for c in torch.unique(cls_id).tolist():
            pos = torch.nonzero(cls_id == c).flatten()  # ascending rank = descending score
            boxes, conf = img[pos, :4], img[pos, 4:5]
            live = torch.ones(pos.numel(), dtype=torch.bool)
            while True:
                nz = torch.nonzero(live).flatten()
                if nz.numel() == 0:
                    break
                top = nz[0]
                hit = (bbox_iou_plus1(boxes[top].unsqueeze(0), boxes[nz]) > nms_thres)
                members = nz[hit]
                w = conf[members]
                merged[int(pos[top])] = (w * boxes[members]).sum(0) / w.sum()
                keep_rank.append(int(pos[top]))
                live[members] = False
                live[top] = False  # reference would spin forever if the top row does not hit itself
        keep_rank.sort()
        kr = torch.tensor(keep_rank, dtype=torch.long)
        det = torch.cat((torch.stack([merged[r] for r in keep_rank]), img[kr, 4:5],
                         cls_conf[kr].unsqueeze(1), cls_id[kr].float().unsqueeze(1)), 1)
        out[bi], out_idx[bi] = det, cand[kr]
    return out, out_idx


# --------------------------------------------------------------------------------------------
# D4 + N2 — SSD / RetinaNet prior decode and top-k class-agnostic NMS
# --------------------------------------------------------------------------------------------

def prior_decode(loc: torch.Tensor, priors: torch.Tensor) -> torch.Tensor:
    """loc[P,4] offsets + priors[P,4] (cx,cy,w,h) -> corner boxes [P,4].  model/SSD.py:253-258."""
    xy = loc[:, :2] * priors[:, 2:] + priors[:, :2]
    wh = loc[:, 2:].exp() * priors[:, 2:]
    return torch.cat([xy - wh / 2, xy + wh / 2], 1)


def ssd_nms(loc_preds: torch.Tensor, cls_preds: torch.Tensor, priors: torch.Tensor, topk: int = 100,
            nms_thresh: float = 0.5, class_thresh: float = 0.45, mode: str = "union",
            return_index: bool = False):
    """SSD / RetinaNet post-processing, including its quirks.  model/SSD.py:249-310 ≡ RetinaNet.py:117-178.

    Quirks reproduced: (i) the loop stops when one box is left WITHOUT keeping it
    (SSD.py:277-278); (ii) `keep` holds indices into the score-filtered set but indexes the UNFILTERED
    `boxes` / `labels` (SSD.py:303-307); (iii) exactly one candidate raises IndexError (SSD.py:262,266).
    `return_index` additionally returns the filtered-space indices (== rows of boxes/labels used).
    """
    if mode not in ("union", "min"):
        raise TypeError("Unknown nms mode: %s." % mode)  # SSD.py:298-299
    out, out_idx = [], []
    for bi in range(loc_preds.shape[0]):
        boxes = prior_decode(loc_preds[bi], priors)
        score, labels = cls_preds[bi].sigmoid().max(1)  # SSD.py:260
        sel = torch.nonzero(score > class_thresh).flatten()  # SSD.py:261-262 (squeeze handled below)
        if sel.numel() == 1:
            raise IndexError("too many indices for tensor of dimension 1")  # SSD.py:266 on a 0-dim index
        fb, fs = boxes[sel], score[sel]
        x1, y1, x2, y2 = fb[:, 0], fb[:, 1], fb[:, 2], fb[:, 3]
        areas = (x2 - x1 + 1) * (y2 - y1 + 1)  # SSD.py:271
        order = torch.sort(fs, stable=True, dim=0, descending=True)[1][:topk]  # SSD.py:272-273
        keep = []
        while order.numel() > 1:  # SSD.py:276-278 (a single survivor is dropped)
            i = int(order[0])
            keep.append(i)
            rest = order[1:]
            xx1 = x1[rest].clamp(min=float(x1[i]))  # SSD.py:282-285
            yy1 = y1[rest].clamp(min=float(y1[i]))
            xx2 = x2[rest].clamp(max=float(x2[i]))
            yy2 = y2[rest].clamp(max=float(y2[i]))
            inter = (xx2 - xx1 + 1).clamp(min=0) * (yy2 - yy1 + 1).clamp(min=0)  # SSD.py:287-289
            if mode == "union":
                ovr = inter / (areas[i] + areas[rest] - inter)  # SSD.py:292
            else:
                ovr = inter / areas[rest].clamp(max=float(areas[i]))  # SSD.py:294
            ok = torch.nonzero(ovr <= nms_thresh).flatten()  # SSD.py:298
            if ok.numel() == 0:
                break
            order = rest[ok]  # SSD.py:301
        k = torch.tensor(keep, dtype=torch.long)
        out.append(torch.cat([boxes[k], torch.zeros((len(keep), 1), dtype=F32), fs[k].unsqueeze(1),
                              labels[k].unsqueeze(1).to(F32)], dim=1))  # SSD.py:303-308
        out_idx.append(k)
    return (out, out_idx) if return_index else out


# --------------------------------------------------------------------------------------------
# D1 / D2 / D3 — YOLO head decode formulas (full map)
# --------------------------------------------------------------------------------------------

def _planar5(head: torch.Tensor, num_anchors: int):
    b, g = head.shape[0], head.shape[2]
    fields = head.numel() // (b * num_anchors * g * g)
    return head.reshape(b, num_anchors, fields, g, g).permute(0, 1, 3, 4, 2).contiguous(), g


def decode_yolo_exp(head: torch.Tensor, scaled_anchors: torch.Tensor, stride: float) -> torch.Tensor:
    """D1 (v2-v4 test/loss-time decode) -> [B, A*G*G, 5+C]: xywh*stride, sigmoid conf/cls.
    LightningFunc/accuracy.py:402-435,459-466 ≡ losses.py:679-703."""
    a = scaled_anchors.shape[0]
    p, g = _planar5(head, a)
    gx = torch.arange(g, dtype=F32).view(1, 1, 1, g)  # accuracy.py:424 (x varies along the last dim)
    gy = torch.arange(g, dtype=F32).view(1, 1, g, 1)  # accuracy.py:425
    aw = scaled_anchors[:, 0].view(1, a, 1, 1)
    ah = scaled_anchors[:, 1].view(1, a, 1, 1)
    box = torch.stack((torch.sigmoid(p[..., 0]) + gx, torch.sigmoid(p[..., 1]) + gy,
                       torch.exp(p[..., 2]) * aw, torch.exp(p[..., 3]) * ah), -1)  # accuracy.py:432-435
    b = p.shape[0]
    return torch.cat((box.view(b, -1, 4) * stride, torch.sigmoid(p[..., 4]).view(b, -1, 1),
                      torch.sigmoid(p[..., 5:]).view(b, -1, p.shape[-1] - 5)), -1)  # accuracy.py:459-466


def decode_yolov5(head: torch.Tensor, anchor_grid: torch.Tensor, stride: float) -> torch.Tensor:
    """D2 full-map form -> [B, A*G*G, 5+C]: y=sigmoid(x); xy=(y*2-0.5+grid)*stride; wh=(y*2)^2*anchor.
    LightningFunc/utils/YoloV5Utils.py:241-248 (≡ matched-rows form losses.py:115-116).
    `anchor_grid[A,2]` are PIXEL anchors."""
    a = anchor_grid.shape[0]
    p, g = _planar5(head, a)
    y = p.sigmoid()
    gx = torch.arange(g, dtype=F32).view(1, 1, 1, g)
    gy = torch.arange(g, dtype=F32).view(1, 1, g, 1)
    grid = torch.stack((gx.expand(1, 1, g, g), gy.expand(1, 1, g, g)), -1)
    y[..., 0:2] = (y[..., 0:2] * 2. - 0.5 + grid) * stride
    y[..., 2:4] = (y[..., 2:4] * 2) ** 2 * anchor_grid.view(1, a, 1, 1, 2)
    return y.view(p.shape[0], -1, p.shape[-1])


def decode_yolov4_norm(head: torch.Tensor, anchors: torch.Tensor, scale_x_y: float = 1.0):
    """D3 -> (boxes [B,N,4] normalised corner boxes, confs [B,N,C] = sigmoid(cls)*sigmoid(obj)).
    LightningFunc/utils/YoloV4Utils.py:36-176 (`anchors[A,2]` in grid units, as passed there)."""
    a = anchors.shape[0]
    p, g = _planar5(head, a)  # [B,A,H,W,5+C]
    h, w = p.shape[2], p.shape[3]
    bxy = torch.sigmoid(p[..., 0:2]) * scale_x_y - 0.5 * (scale_x_y - 1)  # YoloV4Utils.py:84
    bwh = torch.exp(p[..., 2:4])
    gx = torch.arange(w, dtype=F32).view(1, 1, 1, w)
    gy = torch.arange(h, dtype=F32).view(1, 1, h, 1)
    bx = (bxy[..., 0] + gx) / w  # YoloV4Utils.py:117,147
    by = (bxy[..., 1] + gy) / h
    bw = bwh[..., 0] * anchors[:, 0].view(1, a, 1, 1) / w  # YoloV4Utils.py:121,147
    bh = bwh[..., 1] * anchors[:, 1].view(1, a, 1, 1) / h
    x1 = bx - bw * 0.5  # YoloV4Utils.py:156-159
    y1 = by - bh * 0.5
    boxes = torch.stack((x1, y1, x1 + bw, y1 + bh), -1).view(p.shape[0], -1, 4)
    confs = (torch.sigmoid(p[..., 5:]) * torch.sigmoid(p[..., 4:5])).view(p.shape[0], -1, p.shape[-1] - 5)
    return boxes, confs


# --------------------------------------------------------------------------------------------
# T1 — build_targets (v2..v4 grid scatter builder)
# --------------------------------------------------------------------------------------------

def build_targets(pred_boxes: torch.Tensor, pred_cls: torch.Tensor, target: torch.Tensor,
                  anchors: torch.Tensor, ignore_thres: float):
    """LightningFunc/accuracy.py:305-380.  Returns the 10-tuple with the reference dtypes
    (masks uint8).  Duplicate (b,a,gj,gi) cells: the last target row wins (CPU index_put_ order),
    `tcls` becomes multi-hot.  The all-or-nothing bounds guards of :340-344 / :361-367 are kept."""
    nB, nA, nG = pred_boxes.shape[0], pred_boxes.shape[1], pred_boxes.shape[2]
    nC = pred_cls.shape[-1]
    obj = torch.zeros(nB, nA, nG, nG, dtype=torch.uint8)
    noobj = torch.ones(nB, nA, nG, nG, dtype=torch.uint8)
    class_mask = torch.zeros(nB, nA, nG, nG, dtype=F32)
    iou_scores = torch.zeros(nB, nA, nG, nG, dtype=F32)
    tx, ty, tw, th = (torch.zeros(nB, nA, nG, nG, dtype=F32) for _ in range(4))
    tcls = torch.zeros(nB, nA, nG, nG, nC, dtype=F32)

    tb = target[:, 2:6] * nG  # accuracy.py:327
    gxy, gwh = tb[:, :2], tb[:, 2:]
    ious = torch.stack([bbox_wh_iou(an, gwh) for an in anchors])  # accuracy.py:331
    best_n = ious.max(0)[1]  # accuracy.py:332
    b = target[:, 0].long()
    lab = target[:, 1].long()
    gx, gy, gw, gh = gxy[:, 0], gxy[:, 1], gwh[:, 0], gwh[:, 1]
    gi, gj = gxy[:, 0].long(), gxy[:, 1].long()  # accuracy.py:337

    idx_ok = not (bool((b >= nB).any()) or bool((best_n >= nA).any()) or bool((gj >= nG).any())
                  or bool((gi >= nG).any()))  # accuracy.py:340-344 (only upper bounds are checked)
    if idx_ok:
        obj[b, best_n, gj, gi] = 1
        noobj[b, best_n, gj, gi] = 0
    for t in range(target.shape[0]):  # accuracy.py:349-358
        if b[t] >= nB or gj[t] >= nG or gi[t] >= nG:
            continue
        noobj[b[t], ious[:, t] > ignore_thres, gj[t], gi[t]] = 0
    if idx_ok and not bool((lab >= nC).any()):  # accuracy.py:361-367
        tx[b, best_n, gj, gi] = gx - gx.floor()
        ty[b, best_n, gj, gi] = gy - gy.floor()
        tw[b, best_n, gj, gi] = torch.log(gw / anchors[best_n][:, 0] + 1e-16)
        th[b, best_n, gj, gi] = torch.log(gh / anchors[best_n][:, 1] + 1e-16)
        tcls[b, best_n, gj, gi, lab] = 1
        class_mask[b, best_n, gj, gi] = (pred_cls[b, best_n, gj, gi].argmax(-1) == lab).float()
        iou_scores[b, best_n, gj, gi] = bbox_iou_plus1(pred_boxes[b, best_n, gj, gi], tb, x1y1x2y2=False)
    return iou_scores, class_mask, obj, noobj, tx, ty, tw, th, tcls, obj.float()


# --------------------------------------------------------------------------------------------
# T2 — build_targets_v5
# --------------------------------------------------------------------------------------------

def build_targets_v5(shapes: Sequence[Sequence[int]], targets: torch.Tensor, anchors: torch.Tensor,
                     nl: int, na: int):
    """LightningFunc/accuracy.py:472-521 ('rect4' style, g=0.5, anchor_t=4.0).

    `shapes[i]` is the shape of level i's prediction [B,na,ny,nx,5+C] (only [2],[3] are used, :483).
    Row order (part of the contract): anchor-major/target-minor kept rows first, then the x-left,
    y-up, x-right, y-down neighbour copies (:505)."""
    nt = targets.shape[0]
    tcls, tbox, indices, anch = [], [], [], []
    gain = torch.ones(6, dtype=F32)
    off = torch.tensor([[1, 0], [0, 1], [-1, 0], [0, -1]], dtype=F32)
    a_idx = torch.arange(na).view(na, 1).repeat(1, nt)  # accuracy.py:477
    for i in range(nl):
        anc = anchors[i]
        ny, nx = shapes[i][2], shapes[i][3]
        gain[2:] = torch.tensor([nx, ny, nx, ny], dtype=F32)  # accuracy.py:483
        t = targets * gain
        a = torch.zeros(0, dtype=torch.long)
        offsets = torch.zeros(0, 2)
        if nt:
            r = t[None, :, 4:6] / anc[:, None]  # accuracy.py:488
            keep = torch.max(r, 1. / r).max(2)[0] < 4.0  # accuracy.py:489
            a, t = a_idx[keep], t.repeat(na, 1, 1)[keep]  # accuracy.py:490
            gxy = t[:, 2:4]
            z = torch.zeros_like(gxy)
            g = 0.5
            j, k = ((gxy % 1. < g) & (gxy > 1.)).T  # accuracy.py:503
            l, m = ((gxy % 1. > (1 - g)) & (gxy < (gain[[2, 3]] - 1.))).T  # accuracy.py:504
            a = torch.cat((a, a[j], a[k], a[l], a[m]), 0)  # accuracy.py:505
            t = torch.cat((t, t[j], t[k], t[l], t[m]), 0)
            offsets = torch.cat((z, z[j] + off[0], z[k] + off[1], z[l] + off[2], z[m] + off[3]), 0) * g
        else:
            t = t[:0]
        bc = t[:, :2].long()  # accuracy.py:509
        gxy, gwh = t[:, 2:4], t[:, 4:6]
        gij = (gxy - offsets).long()  # accuracy.py:512
        indices.append((bc[:, 0], a, gij[:, 1], gij[:, 0]))  # accuracy.py:516 (b, a, gj, gi)
        tbox.append(torch.cat((gxy - gij, gwh), 1))  # accuracy.py:517
        anch.append(anc[a])  # accuracy.py:518
        tcls.append(bc[:, 1])  # accuracy.py:519
    return tcls, tbox, indices, anch


def v5_match_level(pi: torch.Tensor, tbox: torch.Tensor, idx, anch: torch.Tensor):
    """T4 — per-level matched-row decode + GIoU + objectness target.  LightningFunc/losses.py:105-123.
    Returns (giou[m], tobj[B,A,G,G])."""
    b, a, gj, gi = idx
    tobj = torch.zeros_like(pi[..., 0])
    if b.shape[0] == 0:
        return torch.zeros(0), tobj
    ps = pi[b, a, gj, gi]  # losses.py:112
    pxy = ps[:, :2].sigmoid() * 2. - 0.5  # losses.py:115
    pwh = (ps[:, 2:4].sigmoid() * 2) ** 2 * anch  # losses.py:116
    giou = bbox_iou_v5(torch.cat((pxy, pwh), 1).t(), tbox.t(), x1y1x2y2=False, GIoU=True)  # losses.py:118
    tobj[b, a, gj, gi] = giou.detach().clamp(0).type(tobj.dtype)  # losses.py:123 (gr = 1.0)
    return giou, tobj


# --------------------------------------------------------------------------------------------
# T5 — SSD prior matching;  T6 — RetinaNet anchor assignment + encoding
# --------------------------------------------------------------------------------------------

def _center_to_points(c: torch.Tensor) -> torch.Tensor:
    """LightningFunc/losses.py:172-185 (corners clamped to [0,1])."""
    if c.shape[0] == 0:
        return c
    return torch.cat([torch.clamp(c[:, :2] - c[:, 2:] / 2.0, min=0.0),
                      torch.clamp(c[:, :2] + c[:, 2:] / 2.0, max=1.0)], 1)


def ssd_match(default_boxes: torch.Tensor, gt_boxes: torch.Tensor, match_thresh: float = 0.5):
    """LightningFunc/losses.py:199-218.  default_boxes[P,4], gt_boxes[M,4] (cx,cy,w,h normalised).
    Returns (box_with_annotation[P] int64, matched[P] bool)."""
    m = gt_boxes.shape[0]
    d = _center_to_points(default_boxes).unsqueeze(0).expand(m, -1, -1)
    g = _center_to_points(gt_boxes).unsqueeze(1).expand_as(d)
    ious = pair_iou(d, g)  # [M,P]  losses.py:209
    best_prior = ious.max(1)[1]  # losses.py:211
    iou_max, gt_of_prior = ious.max(0)  # losses.py:214
    matched = iou_max >= match_thresh
    matched[best_prior] = True  # losses.py:216
    gt_of_prior[best_prior] = torch.arange(m, dtype=torch.long)  # losses.py:217 (dup: last GT wins)
    return gt_of_prior, matched


def retina_box_iou_xywh(box1: torch.Tensor, box2: torch.Tensor) -> torch.Tensor:
    """[N,M] IoU_+1 (no eps) of centre-size boxes.  LightningFunc/losses.py:361-403."""
    b1 = torch.cat([box1[:, :2] - box1[:, 2:] / 2, box1[:, :2] + box1[:, 2:] / 2], 1)  # losses.py:373
    b2 = torch.cat([box2[:, :2] - box2[:, 2:] / 2, box2[:, :2] + box2[:, 2:] / 2], 1)
    lt = torch.max(b1[:, None, :2], b2[:, :2])  # losses.py:393-394
    rb = torch.min(b1[:, None, 2:], b2[:, 2:])
    wh = (rb - lt + 1).clamp(min=0)  # losses.py:396
    inter = wh[:, :, 0] * wh[:, :, 1]
    a1 = (b1[:, 2] - b1[:, 0] + 1) * (b1[:, 3] - b1[:, 1] + 1)  # losses.py:399-400
    a2 = (b2[:, 2] - b2[:, 0] + 1) * (b2[:, 3] - b2[:, 1] + 1)
    return inter / (a1[:, None] + a2 - inter)  # losses.py:401


def retina_assign(anchors: torch.Tensor, targets: torch.Tensor, batch_size: int, img_size: float):
    """LightningFunc/losses.py:423-445.  anchors[A,4] pixel cxcywh, targets[nt,6] (img,cls,cx,cy,w,h
    normalised).  Returns (loc_targets[B,A,4] fp32, cls_targets[B,A] int64).  Every image must own at
    least one target (the reference's `max(1)` over an empty dim raises otherwise)."""
    locs, clss = [], []
    for bid in range(batch_size):
        sel = targets[:, 0] == bid
        boxes = targets[sel][:, 2:] * img_size  # losses.py:425
        labels = targets[sel][:, 1].long()  # losses.py:426
        ious = retina_box_iou_xywh(anchors, boxes)  # losses.py:427
        max_iou, max_id = ious.max(1)  # losses.py:431
        bx = boxes[max_id]
        loc = torch.cat([(bx[:, :2] - anchors[:, :2]) / anchors[:, 2:],
                         torch.log(bx[:, 2:] / anchors[:, 2:])], 1)  # losses.py:434-436
        cls = 1 + labels[max_id]  # losses.py:437
        cls[max_iou < 0.5] = 0  # losses.py:439
        cls[(max_iou > 0.4) & (max_iou < 0.5)] = -1  # losses.py:440-441
        locs.append(loc)
        clss.append(cls)
    return torch.stack(locs), torch.stack(clss)


# ---------------------------------------------------------------------------------------------------------
# M1 / M2 — detection metrics: true-positive matching and per-class average precision
# ---------------------------------------------------------------------------------------------------------

def get_batch_statistics(outputs, targets: torch.Tensor, iou_threshold: float):
    """Restates LightningFunc/accuracy.py:116-154.

    Per image with detections (`None` entries are skipped, :122-123): walk the rows in their given order; a row whose
    label (last column, :128) occurs among the image's target labels (:146) is compared with ALL target boxes of the
    image through the +1 IoU (:149, accuracy.py:39, corners); the first maximal target (:149 `.max(0)`) is claimed
    when `iou >= iou_threshold` and nobody claimed it before (:150-152).  The walk stops once every target is
    claimed (:142-143) — which changes nothing, later rows could only hit claimed targets.
    Returns `[tp float64 [K], scores (column 4) [K], labels (column 6) [K]]` per image, as numpy like the reference."""
    import numpy as np
    res = []
    for i, out in enumerate(outputs):
        if out is None:
            continue
        boxes, scores, labels = out[:, :4], out[:, 4], out[:, -1]
        tp = np.zeros(boxes.shape[0])
        ann = targets[targets[:, 0] == i][:, 1:]
        if len(ann):
            t_labels, t_boxes = ann[:, 0], ann[:, 1:]
            claimed = []
            for k in range(boxes.shape[0]):
                if len(claimed) == len(ann):
                    break
                if not bool((t_labels == labels[k]).any()):
                    continue
                v, idx = bbox_iou_plus1(boxes[k].unsqueeze(0), t_boxes).max(0)
                if bool(v >= iou_threshold) and int(idx) not in claimed:
                    tp[k] = 1
                    claimed.append(int(idx))
        res.append([tp, scores.cpu().numpy(), labels.cpu().numpy()])
    return res


def compute_ap(recall, precision):
    """Restates accuracy.py:262-287: sentinels, precision envelope from the right, sum of Δrecall · precision."""
    import numpy as np
    mrec = np.concatenate(([0.0], recall, [1.0]))
    mpre = np.concatenate(([0.0], precision, [0.0]))
    mpre = np.maximum.accumulate(mpre[::-1])[::-1]                  # :277-278 (right-to-left running maximum)
    step = np.where(mrec[1:] != mrec[:-1])[0]                       # :282
    return np.sum((mrec[step + 1] - mrec[step]) * mpre[step + 1])   # :285


def ap_per_class(tp, conf, pred_cls, target_cls):
    """Restates accuracy.py:207-260.  Detections ordered by descending `conf` (:221; numpy's default argsort is
    unstable — ties are ordered by ascending position here); per class of `target_cls` (:225): cumulative TP / FP
    in that order, recall = tpc / (n_gt + 1e-16), precision = tpc / (tpc + fpc), their last values, AP by
    `compute_ap`; a class without predictions scores 0 (:238-241).  Returns p, r, ap, f1 (float64), classes int32."""
    import numpy as np
    tp, conf, pred_cls = np.asarray(tp), np.asarray(conf), np.asarray(pred_cls)
    order = np.argsort(-conf, kind="stable")
    tp, conf, pred_cls = tp[order], conf[order], pred_cls[order]
    classes = np.unique(target_cls)
    ap, p, r = [], [], []
    for c in classes:
        sel = pred_cls == c
        n_gt = (np.asarray(target_cls) == c).sum()
        n_p = sel.sum()
        if n_p == 0 and n_gt == 0:
            continue
        if n_p == 0 or n_gt == 0:
            ap.append(0); r.append(0); p.append(0)
            continue
        fpc = (1 - tp[sel]).cumsum()
        tpc = tp[sel].cumsum()
        recall = tpc / (n_gt + 1e-16)
        precision = tpc / (tpc + fpc)
        r.append(recall[-1]); p.append(precision[-1])
        ap.append(compute_ap(recall, precision))
    p, r, ap = np.array(p), np.array(r), np.array(ap)
    f1 = 2 * p * r / (p + r + 1e-16)
    return p, r, ap, f1, classes.astype("int32")


def get_yolo_statistics(selfobj, output, target: torch.Tensor):
    """Restates LightningFunc/accuracy.py:382-470 (test-time statistics of YOLOv2..v4, one entry per level).
    `selfobj` provides anchors / anch_masks / num_classes / img_size / ignore_thres like the Lightning module.
    Per level: D1 decode in grid units (:402-435), `build_targets` on it (:437-443), then
      cls_acc = 100·mean(class_mask[obj]), conf_obj = mean(conf[obj]), conf_noobj = mean(conf[noobj]),
      precision = Σ iou50·det / (Σ conf50 + 1e-16), recall50/75 = Σ iou50/75·det / (Σ obj + 1e-16),
      det = conf50 · class_mask · tconf (:447-457), and `output` = the decoded map with boxes · stride (:459-466).
    Returns {grid: [cls_acc, recall50, recall75, precision, conf_obj, conf_noobj (numpy 0-dim), output tensor]}."""
    res = {}
    if not isinstance(output, list):
        output = [output]
    for i, head in enumerate(output):
        if selfobj.anch_masks is not None:
            anchors = [selfobj.anchors[m] for m in selfobj.anch_masks[i]]
        elif len(selfobj.anchors) == 3:
            anchors = selfobj.anchors[i]
        else:
            anchors = selfobj.anchors
        g = head.size(2)
        stride = selfobj.img_size / g
        scaled = torch.tensor([(aw / stride, ah / stride) for aw, ah in anchors], dtype=F32)
        dec = decode_yolo_exp(head, scaled, 1.0)                       # grid units: stride applied at the end (:461)
        b, a, c = head.shape[0], len(anchors), selfobj.num_classes
        pred_boxes = dec[..., :4].reshape(b, a, g, g, 4)
        pred_conf = dec[..., 4].reshape(b, a, g, g)
        pred_cls = dec[..., 5:].reshape(b, a, g, g, c)
        iou_scores, class_mask, obj, noobj, _, _, _, _, _, tconf = build_targets(pred_boxes, pred_cls, target, scaled,
                                                                                 selfobj.ignore_thres)
        obj, noobj = obj.bool(), noobj.bool()
        cls_acc = 100 * class_mask[obj].mean()
        conf_obj = pred_conf[obj].mean()
        conf_noobj = pred_conf[noobj].mean()
        conf50 = (pred_conf > 0.5).float()
        iou50 = (iou_scores > 0.5).float()
        iou75 = (iou_scores > 0.75).float()
        det = conf50 * class_mask * tconf
        precision = torch.sum(iou50 * det) / (conf50.sum() + 1e-16)
        recall50 = torch.sum(iou50 * det) / (obj.sum() + 1e-16)
        recall75 = torch.sum(iou75 * det) / (obj.sum() + 1e-16)
        out = torch.cat((dec[..., :4] * stride, dec[..., 4:]), -1)
        res[g] = [cls_acc.numpy(), recall50.numpy(), recall75.numpy(), precision.numpy(), conf_obj.numpy(),
                  conf_noobj.numpy(), out]
    return res


def focal_bce_with_logits(pred: torch.Tensor, true: torch.Tensor, gamma: float = 1.5, alpha: float = 0.25) -> torch.Tensor:
    """FocalLoss(nn.BCEWithLogitsLoss(pos_weight=1, reduction='mean'), gamma) — LightningFunc/losses.py:37-64:
    elementwise BCE-with-logits times alpha_factor · (1 - p_t)^gamma, then the mean."""
    loss = torch.nn.functional.binary_cross_entropy_with_logits(pred, true, reduction="none")
    prob = torch.sigmoid(pred)
    p_t = true * prob + (1 - true) * (1 - prob)                       # :54
    alpha_factor = true * alpha + (1 - true) * (1 - alpha)            # :55
    return (loss * (alpha_factor * (1.0 - p_t) ** gamma)).mean()      # :56-60


def v5_loss(output, target: torch.Tensor, anchors: torch.Tensor, nl: int, na: int, nc: int):
    """Restates MultiScaleRegionLoss_v5.forward (losses.py:98-152; reduction 'mean', label smoothing 0 => cp, cn = 1, 0;
    focal gamma 1.5).  `anchors` = the criterion's `self.anchors` ([nl, na, 2], pixel anchors / stride, :95-96).
    Differentiable through torch autograd; returns the metrics dict."""
    lcls, lbox, lobj = torch.zeros(1), torch.zeros(1), torch.zeros(1)
    tcls, tbox, indices, anch = build_targets_v5([tuple(p.shape) for p in output], target, anchors, nl, na)
    for i, pi in enumerate(output):
        b, a, gj, gi = indices[i]
        tobj = torch.zeros_like(pi[..., 0])
        nb = b.shape[0]
        if nb:
            ps = pi[b, a, gj, gi]                                                          # :111
            pxy = ps[:, :2].sigmoid() * 2. - 0.5                                           # :115
            pwh = (ps[:, 2:4].sigmoid() * 2) ** 2 * anch[i]                                # :116
            giou = bbox_iou_v5(torch.cat((pxy, pwh), 1).t(), tbox[i].t(), x1y1x2y2=False, GIoU=True)   # :118
            lbox = lbox + (1.0 - giou).mean()                                              # :119
            tobj[b, a, gj, gi] = giou.detach().clamp(0).type(tobj.dtype)                   # :123 (gr = 1)
            if nc > 1:
                t = torch.full_like(ps[:, 5:], 0.0)                                        # :128 (cn)
                t[range(nb), tcls[i]] = 1.0                                                # :129 (cp)
                lcls = lcls + focal_bce_with_logits(ps[:, 5:], t)                          # :131
        lobj = lobj + focal_bce_with_logits(pi[..., 4], tobj)                              # :137
    lbox = lbox * 0.05
    lobj = lobj * 1.0
    lcls = lcls * 0.58
    return {"loss": lbox + lobj + lcls, "Localization": lbox, "Classification": lcls, "Conf_obj": lobj}

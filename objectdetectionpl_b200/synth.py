"""Deterministic synthetic inputs for the detection post-processing / target-assignment path.

All tensors are generated on the CPU with `torch.Generator().manual_seed(seed)` in fp32 so the CPU
oracle and the GPU see identical bytes.  Nothing here touches the GPU; callers copy.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import torch

__all__ = ["yolo_planar", "make_tie_free", "raw_logits", "ssd_priors", "retina_priors", "prior_heads",
           "labels", "YOLOV5_ANCHORS", "YOLOV3_ANCHORS", "grids_for"]

# model/YOLOV5.py:106 and model/YOLOV3.py:43 (pixel anchors, listed per level in the model's level order)
YOLOV5_ANCHORS = [[10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119], [116, 90, 156, 198, 373, 326]]
YOLOV3_ANCHORS = [[(116, 90), (156, 198), (373, 326)], [(30, 61), (62, 45), (59, 119)], [(10, 13), (16, 30), (33, 23)]]


def grids_for(model: str, img: int) -> List[int]:
    """Level grid sizes in the order the model's forward() returns them
    (v5: strides 8,16,32 `model/YOLOV5.py:79,154`; v3/v4: strides 32,16,8 `model/YOLOV3.py:255-271`; v2: 32)."""
    if model == "yolov5":
        return [img // 8, img // 16, img // 32]
    if model in ("yolov3", "yolov4"):
        return [img // 32, img // 16, img // 8]
    if model == "yolov2":
        return [img // 32]
    raise ValueError(model)


def _score_of(level_tensors: Sequence[torch.Tensor], A: int) -> torch.Tensor:
    """conf * max(cls) per candidate in the reference's concatenation order -> [B, N]."""
    out = []
    for t in level_tensors:
        B, G = t.shape[0], t.shape[-2]
        p = t.reshape(B, A, -1, G, G)
        out.append((p[:, :, 4] * p[:, :, 5:].max(2)[0]).reshape(B, -1))
    return torch.cat(out, 1)


def make_tie_free(level_tensors: Sequence[torch.Tensor], A: int, max_rounds: int = 8) -> int:
    """Nudge `conf` (plane 4) by one ulp wherever two candidates of one image share a fp32 score, until
    no duplicates remain (the reference's argsort order on ties is implementation-defined, §8d).
    Works in place on the planar tensors; returns the number of nudges applied."""
    nudged = 0
    for _ in range(max_rounds):
        score = _score_of(level_tensors, A)
        srt, order = torch.sort(score, dim=1, stable=True)
        dup = torch.zeros_like(score, dtype=torch.bool)
        dup_sorted = torch.zeros_like(dup)
        dup_sorted[:, 1:] = srt[:, 1:] == srt[:, :-1]
        dup.scatter_(1, order, dup_sorted)
        n = int(dup.sum())
        if n == 0:
            return nudged
        nudged += n
        off = 0
        for t in level_tensors:
            B, G = t.shape[0], t.shape[-2]
            p = t.reshape(B, A, -1, G, G)  # view on the same (contiguous, planar) storage
            m = dup[:, off:off + A * G * G].reshape(B, A, G, G)
            conf = p[:, :, 4]
            conf[m] = torch.nextafter(conf[m], torch.full_like(conf[m], 2.0))
            off += A * G * G
    raise RuntimeError("could not make scores tie-free")


def yolo_planar(B: int, A: int, C: int, grids: Sequence[int], img: int, seed: int,
                conf_mode: str = "uniform", v5_view: bool = False, tie_free: bool = True) -> List[torch.Tensor]:
    """Per level a planar head tensor with storage [B, A, 5+C, G, G] holding *already decoded* values
    (the reference NMS consumes raw head values as boxes, model/YOLOV3.py:289-305):
    cx,cy ~ U(0,img); w,h ~ U(8, img/4) (> -1, so the reference loop terminates); cls ~ U(0,1);
    conf ~ U(0,1) ('uniform', the all-survive regime) or sigmoid(N(-4,2)) ('sparse').
    Returned with the nominal shape [B, A*(5+C), G, G] (v2-v4) or [B, A, G, G, 5+C] viewed on the same
    planar storage (v5, model/YOLOV5.py:178-183)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for G in grids:
        t = torch.empty(B, A, 5 + C, G, G, dtype=torch.float32)
        t[:, :, 0:2] = torch.rand(B, A, 2, G, G, generator=g) * img
        t[:, :, 2:4] = 8.0 + torch.rand(B, A, 2, G, G, generator=g) * (img / 4 - 8.0)
        if conf_mode == "uniform":
            t[:, :, 4] = torch.rand(B, A, G, G, generator=g)
        elif conf_mode == "sparse":
            t[:, :, 4] = torch.sigmoid(torch.randn(B, A, G, G, generator=g) * 2.0 - 4.0)
        else:
            raise ValueError(conf_mode)
        t[:, :, 5:] = torch.rand(B, A, C, G, G, generator=g)
        out.append(t)
    if tie_free:
        make_tie_free(out, A)
    if v5_view:
        return [t.view(B, A, t.shape[3], t.shape[4], 5 + C) for t in out]
    return [t.view(B, A * (5 + C), t.shape[3], t.shape[4]) for t in out]


def yolo_crowd(B: int, A: int, C: int, grids: Sequence[int], img: int, seed: int, blobs: int = 200,
               keep_frac: float = 0.1) -> List[torch.Tensor]:
    """Dense-crowd stress input (BASELINE config 5): box centres drawn from `blobs` Gaussian clusters
    (sigma 20 px), w ~ U(10,40), h ~ U(30,120); conf such that ~keep_frac of candidates exceed 0.001."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for G in grids:
        t = torch.empty(B, A, 5 + C, G, G, dtype=torch.float32)
        centres = torch.rand(B, blobs, 2, generator=g) * img
        pick = torch.randint(0, blobs, (B, A * G * G), generator=g)
        c = torch.gather(centres, 1, pick.unsqueeze(-1).expand(-1, -1, 2)).view(B, A, G, G, 2)
        c = c + torch.randn(B, A, G, G, 2, generator=g) * 20.0
        t[:, :, 0] = c[..., 0]
        t[:, :, 1] = c[..., 1]
        t[:, :, 2] = 10.0 + torch.rand(B, A, G, G, generator=g) * 30.0
        t[:, :, 3] = 30.0 + torch.rand(B, A, G, G, generator=g) * 90.0
        u = torch.rand(B, A, G, G, generator=g)
        hi = 0.001 + torch.rand(B, A, G, G, generator=g) * 0.999
        lo = torch.rand(B, A, G, G, generator=g) * 0.0009
        t[:, :, 4] = torch.where(u < keep_frac, hi, lo)
        t[:, :, 5:] = torch.rand(B, A, C, G, G, generator=g)
        out.append(t)
    make_tie_free(out, A)
    return [t.view(B, A * (5 + C), t.shape[3], t.shape[4]) for t in out]


def raw_logits(B: int, A: int, C: int, G: int, seed: int) -> torch.Tensor:
    """Raw (undecoded) head logits [B, A*(5+C), G, G] for the decode modes D1-D3:
    t_xy ~ N(0,1), t_wh ~ N(0,0.5), t_obj ~ N(-4,2), t_cls ~ N(-2,1.5)."""
    g = torch.Generator().manual_seed(seed)
    t = torch.empty(B, A, 5 + C, G, G, dtype=torch.float32)
    t[:, :, 0:2] = torch.randn(B, A, 2, G, G, generator=g)
    t[:, :, 2:4] = torch.randn(B, A, 2, G, G, generator=g) * 0.5
    t[:, :, 4] = torch.randn(B, A, G, G, generator=g) * 2.0 - 4.0
    t[:, :, 5:] = torch.randn(B, A, C, G, G, generator=g) * 1.5 - 2.0
    return t.view(B, A * (5 + C), G, G)


def ssd_priors() -> torch.Tensor:
    """The 8732 SSD300 default boxes [P,4] (cx,cy,w,h in [0,1]) — same construction as the reference's
    generator (LightningFunc/utils/SSDUtils.py:5-27: 6 feature maps 38..1, scales 0.07..0.9 rounded to
    2 dp, extra sqrt(s_k*s_k+1) box first, then aspect ratios 1,2,1/2[,3,1/3]; clamp max 1)."""
    fks, nbox = [38, 19, 10, 5, 3, 1], [3, 5, 5, 5, 3, 3]
    ars = [1.0, 2.0, 0.5, 3.0, 1.0 / 3.0]
    m = len(fks)
    sks = [round(0.07 + ((0.9 - 0.07) / (m - 1)) * k, 2) for k in range(m)]
    rows = []
    for k, fk in enumerate(fks):
        for i in range(fk):
            for j in range(fk):
                cx, cy = (i + 0.5) / fk, (j + 0.5) / fk
                s_extra = math.sqrt(sks[k] * sks[min(k + 1, m - 1)])
                rows.append([cx, cy, s_extra, s_extra])
                for ar in ars[:nbox[k]]:
                    rows.append([cx, cy, sks[k] * math.sqrt(ar), sks[k] / math.sqrt(ar)])
    return torch.clamp(torch.tensor(rows, dtype=torch.float64).float(), max=1.0)


def retina_priors(img: int = 800) -> torch.Tensor:
    """RetinaNet anchors [A,4] pixel (cx,cy,w,h) — same construction as the reference's generator
    (LightningFunc/utils/RetinaUtils.py:6-31: areas 32^2..512^2 on p3..p7, ratios 1/2,1,2, scales
    2^0,2^(1/3),2^(2/3); 9 anchors per cell, cell centres (i+0.5)*img/fm)."""
    areas = [32.0 * 32, 64.0 * 64, 128.0 * 128, 256.0 * 256, 512.0 * 512]
    ratios = [0.5, 1.0, 2.0]
    scales = [1.0, pow(2, 1 / 3.0), pow(2, 2 / 3.0)]
    out = []
    for lvl, s in enumerate(areas):
        wh = []
        for ar in ratios:
            h = math.sqrt(s / ar)
            w = ar * h
            for sr in scales:
                wh.append([w * sr, h * sr])
        wh = torch.tensor(wh, dtype=torch.float32)  # [9,2]
        fm = math.ceil(img / 2.0 ** (lvl + 3))
        cell = torch.tensor(float(img)) / torch.tensor(float(fm))
        xs = (torch.arange(fm, dtype=torch.float32) + 0.5) * cell
        cy, cx = torch.meshgrid(xs, xs, indexing="ij")
        xy = torch.stack((cx, cy), -1).view(fm, fm, 1, 2).expand(fm, fm, 9, 2)
        out.append(torch.cat([xy, wh.view(1, 1, 9, 2).expand(fm, fm, 9, 2)], 3).reshape(-1, 4))
    return torch.cat(out, 0)


def prior_heads(B: int, P: int, C: int, seed: int, cls_mean: float = -3.0, cls_std: float = 2.0
                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """(loc[B,P,4] ~ N(0,0.2), cls[B,P,C] ~ N(cls_mean, cls_std)) for SSD / RetinaNet post-processing."""
    g = torch.Generator().manual_seed(seed)
    loc = torch.randn(B, P, 4, generator=g) * 0.2
    cls = torch.randn(B, P, C, generator=g) * cls_std + cls_mean
    return loc, cls


def labels(B: int, C: int, seed: int, max_per_image: int = 100, min_per_image: int = 1) -> torch.Tensor:
    """Targets [nt,6] = (image, class, cx, cy, w, h) normalised, concatenated in image order
    (dataset/Coco.py:210-220 collate layout): n ~ U{min..max} per image, xy ~ U(0.05,0.85),
    wh ~ U(0.01,0.31)."""
    g = torch.Generator().manual_seed(seed)
    rows = []
    for b in range(B):
        n = int(torch.randint(min_per_image, max_per_image + 1, (1,), generator=g))
        t = torch.empty(n, 6, dtype=torch.float32)
        t[:, 0] = b
        t[:, 1] = torch.randint(0, C, (n,), generator=g).float()
        t[:, 2:4] = 0.05 + torch.rand(n, 2, generator=g) * 0.80
        t[:, 4:6] = 0.01 + torch.rand(n, 2, generator=g) * 0.30
        rows.append(t)
    return torch.cat(rows, 0)

"""TEST INFRASTRUCTURE ONLY — writes tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.gen_golden
Every file stores the seeded inputs, the reference outputs and provenance (torch version).  The
reference has no tests or fixtures of its own, so these files are the pin for
`oracle/ref_port.py` and, through it, for the CUDA path.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

from oracle import ref_harness as rh
from objectdetectionpl_b200 import synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
META = dict(torch=torch.__version__, reference="Leyan529/ObjectDetectionPL @ /root/reference")


def _np(t):
    return t.detach().cpu().numpy()


def _save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    arrays["_meta"] = np.array(repr(META))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
    print("wrote", name, {k: getattr(v, "shape", None) for k, v in arrays.items() if k != "_meta"})


def _pack_list(prefix, lst, d):
    d[prefix + "_n"] = np.array(len(lst))
    for i, t in enumerate(lst):
        d[f"{prefix}_{i}"] = np.zeros((0,), np.float32) if t is None else _np(t)
        d[f"{prefix}_{i}_none"] = np.array(t is None)


def yolo_cases():
    cases = [
        # name, version, A, C, grids, img, seed, conf_lo
        ("yolo_v5_tiny", 5, 3, 4, [8, 4, 2], 64, 11, 0.0),
        ("yolo_v5_mid", 5, 3, 20, [20, 10, 5], 160, 12, 0.0),
        ("yolo_v3_mid", 3, 3, 7, [5, 10, 20], 160, 13, 0.0),
        ("yolo_v2_g13", 2, 5, 20, [13], 416, 14, 0.0),
        ("yolo_v4_filter", 4, 3, 3, [6, 12], 96, 15, -0.1),   # some conf < -0.0151 -> filtered rows
    ]
    for name, ver, A, C, grids, img, seed, conf_lo in cases:
        B = 2
        lv = synth.yolo_planar(B, A, C, grids, img, seed, v5_view=(ver == 5), tie_free=False)
        if conf_lo < 0:
            g = torch.Generator().manual_seed(seed + 1000)
            for t in lv:
                p = t.view(B, A, 5 + C, t.shape[2], t.shape[3])
                p[:, :, 4] = conf_lo + torch.rand(p[:, :, 4].shape, generator=g) * (1.0 - conf_lo)
        synth.make_tie_free(lv, A)
        fn = rh.yolo_nms(ver)
        arg = [t.clone() for t in lv]
        if ver != 5 and len(arg) == 1:
            arg = arg[0]  # exercises the "single tensor" branch (YOLOV3.py:281-282)
        out = fn(None, arg)
        d = {f"level_{i}": _np(t) for i, t in enumerate(lv)}
        d["n_levels"] = np.array(len(lv))
        d["A"], d["C"], d["version"] = np.array(A), np.array(C), np.array(ver)
        _pack_list("out", out, d)
        _save(name, **d)


def ssd_cases():
    acc_priors = rh.ref_import("LightningFunc.utils.SSDUtils").get_dboxes()
    ret_priors = rh.ref_import("LightningFunc.utils.RetinaUtils").get_anchor_boxes(torch.Tensor([128, 128]))
    for name, which, priors, C, seed, mean in [("ssd300_c5", "SSD", acc_priors, 5, 21, -3.0),
                                               ("retina128_c6", "RetinaNet", ret_priors, 6, 22, -2.0),
                                               ("ssd300_sparse", "SSD", acc_priors, 3, 23, -6.5)]:
        loc, cls = synth.prior_heads(2, priors.shape[0], C, seed, cls_mean=mean)
        fn = rh.ssd_nms(which)
        out = fn(types.SimpleNamespace(iou_boxes=priors), (loc.clone(), cls.clone()))
        d = dict(loc=_np(loc), cls=_np(cls), priors=_np(priors))
        _pack_list("out", out, d)
        # 'min' mode and other thresholds on the same inputs
        out2 = fn(types.SimpleNamespace(iou_boxes=priors), (loc.clone(), cls.clone()), topk=50, nms_thresh=0.3,
                  class_thresh=0.3, mode="min")
        _pack_list("out_min", out2, d)
        _save(name, **d)
    _save("priors", ssd=_np(acc_priors),
          retina600=_np(rh.ref_import("LightningFunc.utils.RetinaUtils").get_anchor_boxes(torch.Tensor([600, 600]))),
          retina800_head=_np(rh.ref_import("LightningFunc.utils.RetinaUtils").get_anchor_boxes(torch.Tensor([800, 800]))[:4096]),
          retina800_count=np.array(rh.ref_import("LightningFunc.utils.RetinaUtils").get_anchor_boxes(torch.Tensor([800, 800])).shape[0]))


def iou_cases():
    acc = rh.accuracy()
    g = torch.Generator().manual_seed(31)
    n = 257
    xy = torch.rand(n, 2, generator=g) * 100
    wh = torch.rand(n, 2, generator=g) * 40 + 0.5
    b1 = torch.cat([xy, wh], 1)
    b2 = torch.cat([xy + torch.randn(n, 2, generator=g) * 10, wh * (0.5 + torch.rand(n, 2, generator=g))], 1)
    d = dict(b1=_np(b1), b2=_np(b2))
    d["plus1_xywh"] = _np(acc.bbox_iou(b1, b2, x1y1x2y2=False))
    c1, c2 = acc.xywh2xyxy(b1), acc.xywh2xyxy(b2)
    d["c1"], d["c2"] = _np(c1), _np(c2)
    d["plus1_xyxy"] = _np(acc.bbox_iou(c1, c2, x1y1x2y2=True))
    d["plus1_one_vs_all"] = _np(acc.bbox_iou(c1[:1], c2, x1y1x2y2=True))
    for kind in ("IoU", "GIoU", "DIoU", "CIoU"):
        for corner in (False, True):
            a = (c1 if corner else b1).t().clone().requires_grad_(True)
            bb = (c2 if corner else b2).t().clone()
            kw = {} if kind == "IoU" else {kind: True}
            v = acc.bbox_iou_v5(a, bb, x1y1x2y2=corner, **kw)
            gw = torch.linspace(0.5, 1.5, n)
            (v * gw).sum().backward()
            tag = f"v5_{kind}_{'xyxy' if corner else 'xywh'}"
            d[tag], d[tag + "_grad"] = _np(v), _np(a.grad)
    d["pair_iou"] = _np(acc.iou(c1.clamp(0, 100), c2.clamp(0, 100)))
    _save("iou", **d)


def build_targets_cases():
    acc = rh.accuracy()
    for name, B, A, G, C, seed, maxn in [("bt_g13", 4, 3, 13, 5, 41, 12), ("bt_g26_dups", 2, 3, 26, 3, 42, 60)]:
        g = torch.Generator().manual_seed(seed)
        tg = synth.labels(B, C, seed, max_per_image=maxn)
        if "dups" in name:  # force duplicate cells: repeat a block of rows with other classes / sizes
            extra = tg[:10].clone()
            extra[:, 1] = (extra[:, 1] + 1) % C
            extra[:, 4:6] *= 1.1
            tg = torch.cat([tg, extra], 0)
        pred_boxes = torch.rand(B, A, G, G, 4, generator=g) * G
        pred_cls = torch.rand(B, A, G, G, C, generator=g)
        anchors = torch.tensor([[1.25, 1.625], [2.0, 3.75], [4.125, 2.875]])
        out = acc.build_targets(pred_boxes, pred_cls, tg, anchors, 0.5)
        names = ["iou_scores", "class_mask", "obj_mask", "noobj_mask", "tx", "ty", "tw", "th", "tcls", "tconf"]
        d = dict(pred_boxes=_np(pred_boxes), pred_cls=_np(pred_cls), target=_np(tg), anchors=_np(anchors))
        for k, v in zip(names, out):
            d[k] = _np(v)
        _save(name, **d)


def build_targets_v5_cases():
    acc, losses = rh.accuracy(), rh.losses()
    for name, B, C, img, seed, maxn in [("btv5_small", 4, 5, 160, 51, 12), ("btv5_mid", 8, 80, 640, 52, 40)]:
        tg = synth.labels(B, C, seed, max_per_image=maxn)
        stride = torch.tensor([8., 16., 32.])
        anchors = torch.tensor(synth.YOLOV5_ANCHORS).float().view(3, -1, 2) / stride.view(-1, 1, 1)
        g = torch.Generator().manual_seed(seed + 1)
        p = [torch.randn(B, 3, img // s, img // s, 5 + C, generator=g) for s in (8, 16, 32)]
        tcls, tbox, indices, anch = acc.build_targets_v5(p, tg, anchors, 3, 3)
        d = dict(target=_np(tg), anchors=_np(anchors), B=np.array(B), C=np.array(C), img=np.array(img))
        for i in range(3):
            d[f"p_{i}"] = _np(p[i]) if name == "btv5_small" else np.array(p[i].shape)
            d[f"tcls_{i}"], d[f"tbox_{i}"], d[f"anch_{i}"] = _np(tcls[i]), _np(tbox[i]), _np(anch[i])
            for k, nm in enumerate("b a gj gi".split()):
                d[f"{nm}_{i}"] = _np(indices[i][k])
        if name == "btv5_small":
            # reference criterion forward: pins the matched-row decode + GIoU (losses.py:105-123)
            crit = losses.MultiScaleRegionLoss_v5(synth.YOLOV5_ANCHORS, None, None, None, None, C, img)
            pr = [t.clone().requires_grad_(True) for t in p]
            m = crit(pr, tg)
            d["lbox"], d["lobj"], d["lcls"] = _np(m["Localization"]), _np(m["Conf_obj"]), _np(m["Classification"])
            m["Localization"].sum().backward()
            for i in range(3):
                d[f"lbox_grad_{i}"] = _np(pr[i].grad)
        _save(name, **d)


def decode_cases():
    acc = rh.accuracy()
    B, A, C, G, img = 2, 3, 4, 13, 416
    head = synth.raw_logits(B, A, C, G, 61)
    anchors = [(116 / 32, 90 / 32), (156 / 32, 198 / 32), (373 / 32, 326 / 32)]  # YOLOV3.py:55-56 (pre-divided)
    tg = synth.labels(B, C, 62, max_per_image=6)
    selfobj = types.SimpleNamespace(anch_masks=None, anchors=[anchors, anchors, anchors], num_classes=C,
                                    img_size=img, ignore_thres=0.5)
    bm = acc.get_yolo_statistics(selfobj, [head.clone()], tg)
    d = dict(head=_np(head), anchors=np.array(anchors, np.float32), img=np.array(img), target=_np(tg))
    d["d1_output"] = _np(bm[G][6])
    d["d1_scaled_anchors"] = _np(selfobj.scaled_anchors)
    d["d1_metrics"] = np.array([float(x) for x in bm[G][:6]], np.float64)
    v4 = rh.ref_import("LightningFunc.utils.YoloV4Utils")
    flat = [v for a in anchors for v in a]
    boxes, confs = v4.yolo_forward_dynamic(head.clone(), 0.5, C, flat, A, scale_x_y=1.05)
    d["d3_boxes"], d["d3_confs"] = _np(boxes), _np(confs)
    _save("decode", **d)


def match_cases():
    losses = rh.losses()
    g = torch.Generator().manual_seed(71)
    priors = rh.ref_import("LightningFunc.utils.SSDUtils").get_dboxes()
    gt = torch.cat([0.1 + torch.rand(7, 2, generator=g) * 0.8, 0.05 + torch.rand(7, 2, generator=g) * 0.4], 1)
    S = losses.SSDLoss
    ns = types.SimpleNamespace()
    ns.center_to_points = lambda t: S.center_to_points(ns, t)
    ns.expand_defaults_and_annotations = lambda a, b: S.expand_defaults_and_annotations(ns, a, b)
    idx, matched = S.match(ns, priors, gt, 0.5)
    d = dict(priors=_np(priors), gt=_np(gt), ssd_idx=_np(idx), ssd_matched=_np(matched))

    # RetinaNet: capture the targets the criterion builds (losses.py:423-445) via stub criteria.
    anchors = rh.ref_import("LightningFunc.utils.RetinaUtils").get_anchor_boxes(torch.Tensor([160, 160]))
    B, C = 3, 4
    tg = synth.labels(B, C, 72, max_per_image=5)
    tg[:, 4:6] = tg[:, 4:6] * 1.5 + 0.1
    cap = {}

    def cls_crit(num_classes, reduction="sum"):
        def f(pred, tgt):
            cap["cls_t"] = tgt.clone()
            return pred.sum() * 0
        return f

    def coord_crit(reduction="sum"):
        def f(pred, tgt):
            cap["loc_t"] = tgt.clone()
            return pred.sum() * 0
        return f

    crit = losses.RetinaNetLoss(anchors, cls_crit, coord_crit, C, 160)
    A = anchors.shape[0]
    crit((torch.zeros(B, A, 4), torch.zeros(B, A, C)), tg)
    d.update(r_anchors=_np(anchors), r_target=_np(tg), r_B=np.array(B), r_img=np.array(160),
             r_loc_pos=_np(cap["loc_t"]), r_cls_nonignored=_np(cap["cls_t"]))
    d["r_iou"] = _np(crit.box_iou(anchors, tg[tg[:, 0] == 0][:, 2:] * 160, order="xywh"))
    _save("match", **d)


def metrics_cases():
    """get_batch_statistics + ap_per_class of the unmodified reference (accuracy.py:116-154, 207-287) on the output of
    its own NMS; the labels are jittered copies of some detections (so that there are true positives, duplicates aiming
    at the same label, and labels nobody reaches) plus unrelated boxes."""
    acc = rh.accuracy()
    nms = rh.yolo_nms(5)
    for name, B, C, grids, img, seed in (("metrics_small", 3, 4, [10, 5], 80, 81), ("metrics_mid", 6, 12, [20, 10, 5], 160, 82)):
        lv = synth.yolo_planar(B, 3, C, grids, img, seed, v5_view=True)
        dets = nms(None, [t.clone() for t in lv])
        dets[1] = None                                   # an image without detections (accuracy.py:122-123)
        g = torch.Generator().manual_seed(seed)
        tg = []
        for b, d in enumerate(dets):
            if d is None:
                tg.append(torch.tensor([[b, 1.0, 5.0, 5.0, 30.0, 30.0]]))
                continue
            if b == B - 1:
                continue                                 # an image without labels (:133-135)
            k = min(8, d.shape[0])
            pick = d[torch.randperm(d.shape[0], generator=g)[:k]]
            box = pick[:, :4] + torch.randn(k, 4, generator=g) * 2.5
            lab = pick[:, 6:7].clone()
            lab[0] = (lab[0] + 1) % C                    # one label with the wrong class
            rnd = torch.rand(3, 4, generator=g) * img * 0.5
            rnd[:, 2:] += rnd[:, :2] + 4
            tg.append(torch.cat([torch.full((k, 1), float(b)), lab, box], 1))
            tg.append(torch.cat([torch.full((3, 1), float(b)), torch.randint(0, C, (3, 1), generator=g).float(), rnd], 1))
        tg = torch.cat(tg)
        stats = acc.get_batch_statistics(dets, tg, 0.5)
        tp, sc, lb = [np.concatenate(x, 0) for x in zip(*stats)]
        labels = tg[:, 1].tolist()
        p, r, ap, f1, cls = acc.ap_per_class(tp, sc, lb, labels)
        d = dict(targets=_np(tg), thr=np.array(0.5), tp=tp, scores=sc, labels=lb, target_cls=np.array(labels),
                 p=p, r=r, ap=ap, f1=f1, classes=cls, n_stats=np.array(len(stats)))
        _pack_list("dets", dets, d)
        for i, st in enumerate(stats):
            d[f"stat_tp_{i}"] = st[0]
        _save(name, **d)


def yolo_stats_cases():
    """get_yolo_statistics of the unmodified reference (accuracy.py:382-470): a YOLOv3-like three-level call (per-level
    anchor lists), a YOLOv2-like single head (5 anchors) and a YOLOv4-like call through `anch_masks`; the heads are
    shifted so that a good share of the confidences passes 0.5, and labels repeat cells."""
    acc = rh.accuracy()
    C, img, B = 4, 416, 2
    tg = synth.labels(B, C, 94, max_per_image=12)
    tg = torch.cat([tg, tg[:5]])                                     # duplicate rows: same cell, last one wins

    def plant(h, A, G, anchors_px):
        """Make every anchor of the label's cell predict the label (box, class, confidence) for every second label, so
        that the IoU / precision / recall sums are non-trivial."""
        v = h.view(B, A, 5 + C, G, G)
        stride = img / G
        for t in tg[::2]:
            b, lab = int(t[0]), int(t[1])
            gx, gy, gw, gh = (t[2:] * G).tolist()
            gi, gj = int(gx), int(gy)
            if not (0 <= gi < G and 0 <= gj < G):
                continue
            fx = min(max(gx - gi, 1e-3), 1 - 1e-3)
            fy = min(max(gy - gj, 1e-3), 1 - 1e-3)
            for a in range(A):
                aw, ah = anchors_px[a][0] / stride, anchors_px[a][1] / stride
                v[b, a, 0, gj, gi] = float(np.log(fx / (1 - fx)))
                v[b, a, 1, gj, gi] = float(np.log(fy / (1 - fy)))
                v[b, a, 2, gj, gi] = float(np.log(gw / aw))
                v[b, a, 3, gj, gi] = float(np.log(gh / ah))
                v[b, a, 4, gj, gi] = 4.0
                v[b, a, 5:, gj, gi] = -5.0
                v[b, a, 5 + lab, gj, gi] = 5.0
        return h

    a3 = [[(116, 90), (156, 198), (373, 326)], [(30, 61), (62, 45), (59, 119)], [(10, 13), (16, 30), (33, 23)]]
    heads = []
    for lvl, (G, seed) in enumerate(((13, 91), (26, 92), (52, 93))):
        h = synth.raw_logits(B, 3, C, G, seed)
        h.view(B, 3, 5 + C, G, G)[:, :, 4] += 3.5
        heads.append(plant(h, 3, G, a3[lvl]))
    s3 = types.SimpleNamespace(anch_masks=None, anchors=a3, num_classes=C, img_size=img, ignore_thres=0.5)
    bm3 = acc.get_yolo_statistics(s3, [h.clone() for h in heads], tg)
    d = dict(target=_np(tg), img=np.array(img), C=np.array(C))
    for G, h in zip((13, 26, 52), heads):
        d[f"v3_head_{G}"] = _np(h)
        d[f"v3_metrics_{G}"] = np.array([float(x) for x in bm3[G][:6]], np.float64)
        d[f"v3_output_{G}"] = _np(bm3[G][6])
    d["v3_anchors"] = np.array(a3, np.float32)
    # YOLOv2: one head, 5 anchors in one flat list (accuracy.py:396-397)
    a2 = [(1.3221, 1.73145), (3.19275, 4.00944), (5.05587, 8.09892), (9.47112, 4.84053), (11.2364, 10.0071)]
    h2 = synth.raw_logits(B, 5, C, 13, 95)
    h2.view(B, 5, 5 + C, 13, 13)[:, :, 4] += 3.5
    plant(h2, 5, 13, [(w * 32, h_ * 32) for w, h_ in a2])
    s2 = types.SimpleNamespace(anch_masks=None, anchors=a2, num_classes=C, img_size=img, ignore_thres=0.5)
    bm2 = acc.get_yolo_statistics(s2, h2.clone(), tg)
    d["v2_head"], d["v2_anchors"] = _np(h2), np.array(a2, np.float32)
    d["v2_metrics"] = np.array([float(x) for x in bm2[13][:6]], np.float64)
    d["v2_output"] = _np(bm2[13][6])
    # YOLOv4: flat anchor list + masks (accuracy.py:388-389)
    flat = np.array([(12, 16), (19, 36), (40, 28), (36, 75), (76, 55), (72, 146), (142, 110), (192, 243), (459, 401)], np.float32)
    masks = [[6, 7, 8], [3, 4, 5], [0, 1, 2]]
    s4 = types.SimpleNamespace(anch_masks=masks, anchors=flat, num_classes=C, img_size=img, ignore_thres=0.5)
    bm4 = acc.get_yolo_statistics(s4, [h.clone() for h in heads[:2]], tg)
    d["v4_anchors"], d["v4_masks"] = flat, np.array(masks)
    for G in (13, 26):
        d[f"v4_metrics_{G}"] = np.array([float(x) for x in bm4[G][:6]], np.float64)
        d[f"v4_output_{G}"] = _np(bm4[G][6])
    _save("yolo_stats", **d)


def v5_loss_cases():
    """MultiScaleRegionLoss_v5 forward + backward of the unmodified reference (losses.py:68-152): metrics and the
    gradient w.r.t. every head level."""
    losses = rh.losses()
    B, C, img = 2, 4, 160
    anchors_px = [list(map(float, a)) for a in synth.YOLOV5_ANCHORS]
    crit = losses.MultiScaleRegionLoss_v5(anchors_px, None, None, None, None, C, img)
    g = torch.Generator().manual_seed(101)
    p = [torch.randn(B, 3, img // s, img // s, 5 + C, generator=g).requires_grad_(True) for s in (8, 16, 32)]
    tg = synth.labels(B, C, 102, max_per_image=7)
    tg = torch.cat([tg, tg[:3]])                                       # duplicate labels: same cells claimed twice
    m = crit(p, tg)
    m["loss"].backward()
    d = dict(target=_np(tg), C=np.array(C), img=np.array(img), anchors_scaled=_np(crit.anchors),
             metrics=np.array([float(m[k]) for k in ("loss", "Localization", "Classification", "Conf_obj")], np.float64))
    for i, t in enumerate(p):
        d[f"p_{i}"] = _np(t)
        d[f"grad_{i}"] = _np(t.grad)
    # single-class criterion: the class term is skipped (losses.py:127)
    crit1 = losses.MultiScaleRegionLoss_v5(anchors_px, None, None, None, None, 1, img)
    p1 = [torch.randn(B, 3, img // s, img // s, 6, generator=g).requires_grad_(True) for s in (8, 16, 32)]
    tg1 = tg.clone(); tg1[:, 1] = 0
    m1 = crit1(p1, tg1)
    m1["loss"].backward()
    d["metrics_1c"] = np.array([float(m1[k]) for k in ("loss", "Localization", "Classification", "Conf_obj")], np.float64)
    for i, t in enumerate(p1):
        d[f"p1c_{i}"] = _np(t)
        d[f"grad1c_{i}"] = _np(t.grad)
    d["target_1c"] = _np(tg1)
    _save("v5_loss", **d)


def main():
    if not rh.available():
        sys.exit("reference tree not present; golden vectors can only be generated in the build container")
    torch.manual_seed(0)
    yolo_cases()
    ssd_cases()
    iou_cases()
    build_targets_cases()
    build_targets_v5_cases()
    decode_cases()
    match_cases()
    metrics_cases()
    yolo_stats_cases()
    v5_loss_cases()


if __name__ == "__main__":
    main()

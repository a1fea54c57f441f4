"""Timings of the non-headline BASELINE configurations (development / profiles aid; bench.py is the contract).
Prints one JSON object per configuration: CUDA-event time of the public-API call (median of N), algorithmic
bytes and the implied GB/s where the stage is a streaming one."""
import argparse, json, os, statistics, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import objectdetectionpl_b200 as od
from objectdetectionpl_b200 import synth

DEV = torch.device("cuda:0")


def timed(fn, iters=int(os.environ.get("B200DET_CFG_ITERS", "20")), warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts)


def main():
    out = []
    # cfg 2: YOLOv3 416 COCO, batch 64
    lv = [t.to(DEV) for t in synth.yolo_planar(64, 3, 80, [13, 26, 52], 416, 2, tie_free=False)]
    us = timed(lambda: od.yolo_nms_raw(lv, 3))
    nbytes = sum(t.numel() for t in lv) * 4
    out.append(dict(cfg="2: yolov3 416 C=80 B=64 decode+NMS (device-resident, no host sync)", us=us, img_per_s=64 / (us * 1e-6),
                    head_MB=nbytes / 1e6))
    del lv
    # cfg 5: dense-crowd shard (64 of the 512 images: one GPU's share), 1280x1280, 5 classes, conf_thres 0.001
    lv = [t.to(DEV) for t in synth.yolo_crowd(64, 3, 5, [160, 80, 40], 1280, seed=5)]
    us = timed(lambda: od.yolo_nms_raw(lv, 3, conf_thres=0.001))
    out.append(dict(cfg="5: dense-crowd shard, yolov5l 1280 C=5 B=64/GPU, conf_thres 0.001 (~10k pre-NMS boxes/image), decode+NMS",
                    us=us, img_per_s=64 / (us * 1e-6), head_MB=sum(t.numel() for t in lv) * 4 / 1e6))
    del lv
    # cfg 3: SSD300 / RetinaNet 800, batch 32
    for name, pri in (("3a: SSD300 P=8732", synth.ssd_priors()), ("3b: RetinaNet800 P=120087", synth.retina_priors(800))):
        loc, cls = synth.prior_heads(32, pri.shape[0], 80, 3)
        loc, cls, pr = loc.to(DEV), cls.to(DEV), pri.to(DEV)
        us = timed(lambda: od.prior_nms_raw(loc, cls, pr))
        nbytes = (loc.numel() + cls.numel()) * 4
        out.append(dict(cfg=name + " C=80 B=32 prior decode + top-100 NMS", us=us, img_per_s=32 / (us * 1e-6),
                        head_MB=nbytes / 1e6, GBps_whole_pipeline=nbytes / (us * 1e-6) / 1e9))
        del loc, cls
    # cfg 4: YOLOv5s target assignment, batch 64
    B, C = 64, 80
    tg = synth.labels(B, C, 4).to(DEV)
    stride = torch.tensor([8., 16., 32.])
    anchors = torch.tensor(synth.YOLOV5_ANCHORS).float().view(3, -1, 2) / stride.view(-1, 1, 1)
    shapes = [(B, 3, 640 // s, 640 // s, 5 + C) for s in (8, 16, 32)]
    us_bt = timed(lambda: od.build_targets_v5(shapes, tg, anchors, 3, 3))
    p = [torch.randn(s, device=DEV, requires_grad=True) for s in shapes]
    tcls, tbox, idx, anch = od.build_targets_v5(shapes, tg, anchors, 3, 3)

    def fwd_bwd():
        l = 0
        for i in range(3):
            giou, tobj = od.v5_match_level(p[i], tbox[i], idx[i], anch[i])
            l = l + (1.0 - giou).mean()
        l.backward()
    us_m = timed(fwd_bwd)
    out.append(dict(cfg="4: YOLOv5s build_targets_v5 (3 levels, incl. 1 host sync)", us=us_bt, nt=int(tg.shape[0]),
                    rows=[int(t.shape[0]) for t in tcls]))
    out.append(dict(cfg="4: v5 matched-row GIoU fwd+bwd (3 levels, through autograd)", us=us_m))

    anchors_dev = anchors.to(DEV)                  # the criterion keeps its scaled anchors on the device (losses.py:95-96)

    def loss_fwd_bwd():
        for t in p:
            t.grad = None
        od.v5_loss(p, tg, anchors_dev, 3, 3, C)["loss"].backward()
    us_l = timed(loss_fwd_bwd)
    out.append(dict(cfg="4: fused v5 loss (build_targets_v5 + box/obj/cls terms) fwd+bwd, 3 levels, B=64 C=80", us=us_l,
                    head_MB=sum(t.numel() for t in p) * 4 / 1e6))
    del p
    an3 = torch.tensor([[1.25, 1.625], [2.0, 3.75], [4.125, 2.875]], device=DEV)
    for G in (13, 26, 52):
        pb = torch.rand(B, 3, G, G, 4, device=DEV) * G
        pc = torch.rand(B, 3, G, G, C, device=DEV)
        us = timed(lambda: od.build_targets(pb, pc, tg, an3, 0.5))
        out.append(dict(cfg=f"4': build_targets (v2-v4) G={G} B=64 C=80", us=us))
        del pb, pc
    pri = synth.ssd_priors().to(DEV)
    gt = torch.rand(20, 4, device=DEV) * 0.4 + 0.1
    out.append(dict(cfg="T5: SSDLoss.match P=8732 M=20", us=timed(lambda: od.ssd_match(pri, gt, 0.5))))
    anc = synth.retina_priors(600).to(DEV)
    tg32 = synth.labels(32, C, 6).to(DEV)
    out.append(dict(cfg="T6: RetinaNet assign A=67995 B=32", us=timed(lambda: od.retina_assign(anc, tg32, 32, 600.0))))
    # M1 / M2: metrics on the headline NMS output (64 images, ~19.6 k kept rows each, <= 100 labels per image)
    lv = [t.to(DEV) for t in synth.yolo_planar(64, 3, 80, [80, 40, 20], 640, 1234, v5_view=True, tie_free=False)]
    rows, _, count = od.yolo_nms_raw(lv, 3)
    del lv
    n_pad = rows.shape[1]
    row_start = torch.arange(64, device=DEV, dtype=torch.int64) * n_pad
    tgm = synth.labels(64, 80, 4).to(DEV)
    tgm[:, 2:4] = tgm[:, 2:4] * 640
    tgm[:, 4:6] = tgm[:, 2:4] + tgm[:, 4:6] * 640
    us = timed(lambda: od.batch_statistics_raw(rows, row_start, count, n_pad, tgm, 0.5))
    kept = int(count.sum())
    out.append(dict(cfg="M1: get_batch_statistics on the headline NMS output (device-resident)", us=us, detections=kept,
                    labels=int(tgm.shape[0]), det_per_s=kept / (us * 1e-6)))
    tp = od.batch_statistics_raw(rows, row_start, count, n_pad, tgm, 0.5)
    mask = torch.arange(n_pad, device=DEV)[None, :] < count[:, None]
    tpf, conf, cls = tp[mask], rows[..., 4][mask], rows[..., 6][mask]
    classes = torch.arange(80, device=DEV, dtype=torch.int32)
    n_gt = torch.bincount(tgm[:, 1].long(), minlength=80).int()
    us = timed(lambda: od.ap_per_class_device(tpf, conf, cls, classes, n_gt))
    out.append(dict(cfg="M2: ap_per_class over the same detections, 80 classes (device-resident)", us=us, detections=kept,
                    det_per_s=kept / (us * 1e-6)))
    # M3: get_yolo_statistics, YOLOv3 416 COCO heads, batch 64 (device-resident form, per level)
    import types
    a3 = [[(116, 90), (156, 198), (373, 326)], [(30, 61), (62, 45), (59, 119)], [(10, 13), (16, 30), (33, 23)]]
    tgs = synth.labels(64, 80, 4).to(DEV)
    for lvl, G in enumerate((13, 26, 52)):
        head = synth.raw_logits(64, 3, 80, G, 7).to(DEV)
        stride = 416 / G
        sc = torch.tensor([(w / stride, h / stride) for w, h in a3[lvl]], device=DEV)
        us = timed(lambda: od.yolo_statistics_level(head, sc, stride, tgs, 0.5))
        nbytes = head.numel() * 4
        out.append(dict(cfg=f"M3: get_yolo_statistics level G={G} B=64 C=80 (decode + build_targets + metrics, device-resident)",
                        us=us, head_MB=nbytes / 1e6, GBps_read_plus_write=2 * nbytes / (us * 1e-6) / 1e9))
        del head
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()

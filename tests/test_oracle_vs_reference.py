"""Build-container only (needs /root/reference): the oracle restatement against the LIVE unmodified reference on
fresh seeds, and the drop-in installation on the reference's own classes/modules."""
import types

import pytest
import torch

from objectdetectionpl_b200 import synth
from oracle import ref_port as rp

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def rh():
    from oracle import ref_harness
    return ref_harness


@pytest.mark.parametrize("seed", [201, 202, 203])
def test_yolo_nms_live(rh, seed):
    lv = synth.yolo_planar(2, 3, 5, [10, 5], 80, seed, v5_view=True)
    want = rh.yolo_nms(5)(None, [t.clone() for t in lv])
    got = rp.yolo_nms(lv, num_anchors=3)
    for g, w in zip(got, want):
        assert torch.equal(g, w)
    lv3 = [t.reshape(2, -1, t.shape[2], t.shape[3]) for t in lv]
    want3 = rh.yolo_nms(3)(None, [t.clone() for t in lv3])
    for g, w in zip(got, want3):        # v3 == v5 on the same memory (SURVEY ten-things #2)
        assert torch.equal(g, w)


@pytest.mark.parametrize("seed", [211, 212])
def test_ssd_nms_live(rh, seed):
    pri = synth.ssd_priors()[:2000]
    loc, cls = synth.prior_heads(2, 2000, 4, seed, cls_mean=-2.0)
    want = rh.ssd_nms("SSD")(types.SimpleNamespace(iou_boxes=pri), (loc.clone(), cls.clone()))
    got = rp.ssd_nms(loc, cls, pri)
    for g, w in zip(got, want):
        assert torch.equal(g, w)


@pytest.mark.parametrize("seed", [221, 222])
def test_build_targets_live(rh, seed):
    acc = rh.accuracy()
    g = torch.Generator().manual_seed(seed)
    tg = synth.labels(3, 4, seed, max_per_image=9)
    pb = torch.rand(3, 3, 13, 13, 4, generator=g) * 13
    pc = torch.rand(3, 3, 13, 13, 4, generator=g)
    an = torch.tensor([[1.25, 1.625], [2.0, 3.75], [4.125, 2.875]])
    for a, b in zip(rp.build_targets(pb, pc, tg, an, 0.5), acc.build_targets(pb, pc, tg, an, 0.5)):
        assert a.dtype == b.dtype and torch.equal(a, b)
    stride = torch.tensor([8., 16., 32.])
    anchors = torch.tensor(synth.YOLOV5_ANCHORS).float().view(3, -1, 2) / stride.view(-1, 1, 1)
    p = [torch.zeros(3, 3, 160 // s, 160 // s, 9) for s in (8, 16, 32)]
    w = acc.build_targets_v5(p, tg, anchors, 3, 3)
    o = rp.build_targets_v5([t.shape for t in p], tg, anchors, 3, 3)
    for i in range(3):
        assert torch.equal(o[0][i], w[0][i]) and torch.equal(o[1][i], w[1][i]) and torch.equal(o[3][i], w[3][i])
        assert all(torch.equal(x, y) for x, y in zip(o[2][i], w[2][i]))


def test_install_patches_the_reference_seams(rh):
    import objectdetectionpl_b200 as od
    mods = {v: rh.ref_import(f"model.YOLOV{v}") for v in (2, 3, 4, 5)}
    classes = [getattr(mods[v], f"YOLOv{v}") for v in (2, 3, 4, 5)]
    ssd = rh.ref_import("model.SSD").SSD
    ret = rh.ref_import("model.RetinaNet").RetinaNet
    saved = [(c, c.non_max_suppression) for c in classes + [ssd, ret]]
    losses, acc = rh.losses(), rh.accuracy()
    saved_l = {k: getattr(losses, k) for k in ("build_targets_v5", "bbox_iou_v5", "build_targets", "bbox_iou", "iou")}
    saved_a = {k: getattr(acc, k) for k in ("build_targets_v5", "bbox_iou_v5", "build_targets", "bbox_iou", "xywh2xyxy", "iou",
                                            "get_batch_statistics", "ap_per_class")}
    try:
        od.install(*classes, ssd, ret, losses_module=losses, accuracy_module=acc)
        assert classes[0].non_max_suppression is od.non_max_suppression_v2
        assert all(c.non_max_suppression is od.non_max_suppression for c in classes[1:])
        assert ssd.non_max_suppression is od.prior_non_max_suppression and ret.non_max_suppression is od.prior_non_max_suppression
        assert losses.build_targets_v5 is od.build_targets_v5 and losses.bbox_iou_v5 is od.bbox_iou_v5
        assert acc.xywh2xyxy is od.xywh2xyxy
        assert acc.get_batch_statistics is od.get_batch_statistics and acc.ap_per_class is od.ap_per_class
        # a criterion built after install() picks the replacement up (losses.py:654)
        crit = losses.RegionLoss_v3([(1, 1)] * 3, torch.nn.BCELoss, torch.nn.MSELoss, torch.nn.BCELoss, 4)
        assert crit.build_targets is od.build_targets
    finally:
        for c, f in saved:
            c.non_max_suppression = f
        for k, v in saved_l.items():
            setattr(losses, k, v)
        for k, v in saved_a.items():
            setattr(acc, k, v)


@pytest.mark.parametrize("seed", [231, 232])
def test_metrics_live(rh, seed):
    import numpy as np
    acc = rh.accuracy()
    lv = synth.yolo_planar(3, 3, 5, [10, 5], 80, seed, v5_view=True)
    dets = rp.yolo_nms(lv, num_anchors=3)
    g = torch.Generator().manual_seed(seed)
    tg = []
    for b, d in enumerate(dets):
        pick = d[torch.randperm(d.shape[0], generator=g)[:7]]
        tg.append(torch.cat([torch.full((7, 1), float(b)), pick[:, 6:7], pick[:, :4] + torch.randn(7, 4, generator=g) * 3], 1))
    tg = torch.cat(tg)
    want = acc.get_batch_statistics(dets, tg, 0.5)
    got = rp.get_batch_statistics(dets, tg, 0.5)
    for a, b in zip(got, want):
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    tp, sc, lb = [np.concatenate(x, 0) for x in zip(*want)]
    for a, b in zip(rp.ap_per_class(tp, sc, lb, tg[:, 1].tolist()), acc.ap_per_class(tp, sc, lb, tg[:, 1].tolist())):
        np.testing.assert_allclose(a, b, rtol=1e-12, atol=0)


def test_install_metrics_patches_step_namespace(rh):
    import objectdetectionpl_b200 as od
    acc = rh.accuracy()
    ns = types.SimpleNamespace(get_batch_statistics=acc.get_batch_statistics, ap_per_class=acc.ap_per_class)
    od.install_metrics(step_module=ns)
    assert ns.get_batch_statistics is od.get_batch_statistics and ns.ap_per_class is od.ap_per_class

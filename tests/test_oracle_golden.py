"""CPU: the oracle restatement (oracle/ref_port.py) against the golden vectors produced by the
unmodified reference (tests/golden, oracle/gen_golden.py).  This is the pin for the oracle."""
import numpy as np
import pytest
import torch

from oracle import ref_port as rp
from objectdetectionpl_b200 import synth
from tests.golden_io import load, T, unpack_list, yolo_levels, YOLO_CASES, SSD_CASES, assert_rows_close


@pytest.mark.parametrize("name", YOLO_CASES)
def test_yolo_nms_matches_reference(name):
    d = load(name)
    want = unpack_list(d, "out")
    got = rp.yolo_nms(yolo_levels(d), num_anchors=int(d["A"]))
    assert len(got) == len(want)
    for i, (g, w) in enumerate(zip(got, want)):
        assert_rows_close(g, w, rtol=0, atol=0, what=f"{name}[{i}]")  # same ops, same order: bit-exact


@pytest.mark.parametrize("name", YOLO_CASES)
def test_yolo_nms_fast_equals_loop(name):
    d = load(name)
    rows = rp.yolo_rows_from_planar(yolo_levels(d), int(d["A"]))
    a, ai = rp.yolo_nms_rows(rows, return_index=True)
    b, bi = rp.yolo_nms_fast(rows)
    for i in range(len(a)):
        assert torch.equal(ai[i], bi[i])
        assert_rows_close(b[i], a[i], rtol=1e-6, atol=1e-5, what=f"{name}[{i}]")


@pytest.mark.parametrize("name", SSD_CASES)
def test_ssd_nms_matches_reference(name):
    d = load(name)
    loc, cls, pri = T(d["loc"]), T(d["cls"]), T(d["priors"])
    for got, want in ((rp.ssd_nms(loc, cls, pri), unpack_list(d, "out")),
                      (rp.ssd_nms(loc, cls, pri, topk=50, nms_thresh=0.3, class_thresh=0.3, mode="min"),
                       unpack_list(d, "out_min"))):
        assert len(got) == len(want)
        for i, (g, w) in enumerate(zip(got, want)):
            assert_rows_close(g, w, rtol=0, atol=0, what=f"{name}[{i}]")


def test_ssd_nms_single_candidate_raises():
    pri = synth.ssd_priors()[:16]
    loc = torch.zeros(1, 16, 4)
    cls = torch.full((1, 16, 3), -9.0)
    cls[0, 5, 1] = 4.0
    with pytest.raises(IndexError):
        rp.ssd_nms(loc, cls, pri)
    cls[0, 5, 1] = -9.0
    out = rp.ssd_nms(loc, cls, pri)
    assert out[0].shape == (0, 7)


def test_priors_match_reference_generators():
    d = load("priors")
    assert torch.equal(synth.ssd_priors(), T(d["ssd"]))
    assert torch.equal(synth.retina_priors(600), T(d["retina600"]))
    r800 = synth.retina_priors(800)
    assert r800.shape[0] == int(d["retina800_count"]) == 120087
    assert torch.equal(r800[:4096], T(d["retina800_head"]))


def test_iou_family():
    d = load("iou")
    b1, b2, c1, c2 = T(d["b1"]), T(d["b2"]), T(d["c1"]), T(d["c2"])
    assert torch.equal(rp.xywh2xyxy(b1), c1)
    assert torch.equal(rp.bbox_iou_plus1(b1, b2, x1y1x2y2=False), T(d["plus1_xywh"]))
    assert torch.equal(rp.bbox_iou_plus1(c1, c2), T(d["plus1_xyxy"]))
    assert torch.equal(rp.bbox_iou_plus1(c1[:1], c2), T(d["plus1_one_vs_all"]))
    assert torch.equal(rp.pair_iou(c1.clamp(0, 100), c2.clamp(0, 100)), T(d["pair_iou"]))
    n = b1.shape[0]
    gw = torch.linspace(0.5, 1.5, n)
    for kind in ("IoU", "GIoU", "DIoU", "CIoU"):
        for corner in (False, True):
            a = (c1 if corner else b1).t().clone().requires_grad_(True)
            bb = (c2 if corner else b2).t().clone()
            kw = {} if kind == "IoU" else {kind: True}
            v = rp.bbox_iou_v5(a, bb, x1y1x2y2=corner, **kw)
            (v * gw).sum().backward()
            tag = f"v5_{kind}_{'xyxy' if corner else 'xywh'}"
            assert torch.equal(v.detach(), T(d[tag])), tag
            torch.testing.assert_close(a.grad, T(d[tag + "_grad"]), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("name", ["bt_g13", "bt_g26_dups"])
def test_build_targets(name):
    d = load(name)
    out = rp.build_targets(T(d["pred_boxes"]), T(d["pred_cls"]), T(d["target"]), T(d["anchors"]), 0.5)
    names = ["iou_scores", "class_mask", "obj_mask", "noobj_mask", "tx", "ty", "tw", "th", "tcls", "tconf"]
    for k, v in zip(names, out):
        w = T(d[k])
        assert v.dtype == w.dtype, k
        assert torch.equal(v, w), k


@pytest.mark.parametrize("name", ["btv5_small", "btv5_mid"])
def test_build_targets_v5(name):
    d = load(name)
    shapes = [tuple(d[f"p_{i}"].shape) if name == "btv5_small" else tuple(int(x) for x in d[f"p_{i}"]) for i in range(3)]
    tcls, tbox, idx, anch = rp.build_targets_v5(shapes, T(d["target"]), T(d["anchors"]), 3, 3)
    for i in range(3):
        assert torch.equal(tcls[i], T(d[f"tcls_{i}"]))
        assert torch.equal(tbox[i], T(d[f"tbox_{i}"]))
        assert torch.equal(anch[i], T(d[f"anch_{i}"]))
        for k, nm in enumerate("b a gj gi".split()):
            assert torch.equal(idx[i][k], T(d[f"{nm}_{i}"])), (i, nm)


def test_v5_match_level_pins_lbox():
    d = load("btv5_small")
    tg, anchors = T(d["target"]), T(d["anchors"])
    p = [T(d[f"p_{i}"]).clone().requires_grad_(True) for i in range(3)]
    tcls, tbox, idx, anch = rp.build_targets_v5([t.shape for t in p], tg, anchors, 3, 3)
    lbox = 0
    for i in range(3):
        giou, tobj = rp.v5_match_level(p[i], tbox[i], idx[i], anch[i])
        lbox = lbox + (1.0 - giou).mean()
    lbox = lbox * 0.05
    torch.testing.assert_close(lbox.detach().reshape(1), T(d["lbox"]), rtol=1e-6, atol=1e-7)
    lbox.backward()
    for i in range(3):
        torch.testing.assert_close(p[i].grad, T(d[f"lbox_grad_{i}"]), rtol=1e-5, atol=1e-8)


def test_decode_formulas():
    d = load("decode")
    head = T(d["head"])
    img, G = int(d["img"]), head.shape[2]
    stride = img / G
    scaled = torch.tensor([(aw / stride, ah / stride) for aw, ah in d["anchors"].tolist()])
    assert torch.equal(scaled, T(d["d1_scaled_anchors"]))
    out = rp.decode_yolo_exp(head, scaled, stride)
    assert torch.equal(out, T(d["d1_output"]))
    boxes, confs = rp.decode_yolov4_norm(head, T(d["anchors"]), scale_x_y=1.05)
    torch.testing.assert_close(boxes, T(d["d3_boxes"]).squeeze(2), rtol=1e-6, atol=1e-7)
    # layout-dependent vectorised-vs-scalar sigmoid on the CPU differs in the last ulp
    torch.testing.assert_close(confs, T(d["d3_confs"]), rtol=1e-6, atol=1e-9)


def test_decode_v5_consistent_with_loss_form():
    """D2 has no executable full-map reference (commented out, YoloV5Utils.py:233-250); pin it to the
    matched-row form the loss uses (losses.py:115-116): same sigmoid / *2 / **2 / anchor arithmetic."""
    B, A, C, G = 2, 3, 4, 8
    head = synth.raw_logits(B, A, C, G, 7)
    anchor_grid = torch.tensor([[10., 13.], [16., 30.], [33., 23.]])
    stride = 8.0
    out = rp.decode_yolov5(head, anchor_grid, stride).view(B, A, G, G, 5 + C)
    ps = head.view(B, A, 5 + C, G, G).permute(0, 1, 3, 4, 2)
    pxy = ps[..., :2].sigmoid() * 2. - 0.5
    pwh = (ps[..., 2:4].sigmoid() * 2) ** 2 * (anchor_grid / stride).view(1, A, 1, 1, 2)
    gy, gx = torch.meshgrid(torch.arange(G).float(), torch.arange(G).float(), indexing="ij")
    torch.testing.assert_close(out[..., 0], (pxy[..., 0] + gx) * stride, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(out[..., 1], (pxy[..., 1] + gy) * stride, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(out[..., 2:4], pwh * stride, rtol=1e-6, atol=1e-6)
    assert torch.equal(out[..., 4:], ps[..., 4:].sigmoid())


def test_ssd_match_and_retina_assign():
    d = load("match")
    idx, matched = rp.ssd_match(T(d["priors"]), T(d["gt"]), 0.5)
    assert torch.equal(idx, T(d["ssd_idx"]))
    assert torch.equal(matched, T(d["ssd_matched"]))
    anchors, tg = T(d["r_anchors"]), T(d["r_target"])
    B, img = int(d["r_B"]), float(d["r_img"])
    assert torch.equal(rp.retina_box_iou_xywh(anchors, tg[tg[:, 0] == 0][:, 2:] * img), T(d["r_iou"]))
    loc, cls = rp.retina_assign(anchors, tg, B, img)
    assert torch.equal(loc[cls > 0], T(d["r_loc_pos"]))
    assert torch.equal(cls[cls > -1], T(d["r_cls_nonignored"]))


@pytest.mark.parametrize("name", YOLO_CASES)
def test_c_oracle_matches_reference(name):
    """The plain-C restatement (oracle/c/nms_oracle.c, used for full-size GPU parity) against the reference."""
    from oracle import c_oracle
    d = load(name)
    want = unpack_list(d, "out")
    rows = rp.yolo_rows_from_planar(yolo_levels(d), int(d["A"]))
    got, gidx = c_oracle.yolo_nms_rows(rows)
    _, widx = rp.yolo_nms_rows(rows, return_index=True)
    for i in range(len(want)):
        assert torch.equal(gidx[i], widx[i])
        assert_rows_close(got[i], want[i], rtol=1e-6, atol=1e-4, what=f"{name}[{i}]")


def test_c_oracle_threshold_and_empty():
    from oracle import c_oracle
    lv = synth.yolo_planar(2, 3, 4, [8, 4], 64, 9)
    rows = rp.yolo_rows_from_planar(lv, 3)
    got, _ = c_oracle.yolo_nms_rows(rows, conf_thres=2.0)
    assert got == [None, None]
    a, ai = c_oracle.yolo_nms_rows(rows, conf_thres=0.5, nms_thres=0.6)
    b, bi = rp.yolo_nms_rows(rows, conf_thres=0.5, nms_thres=0.6, return_index=True)
    for i in range(2):
        assert torch.equal(ai[i], bi[i])
        assert_rows_close(a[i], b[i], rtol=1e-6, atol=1e-4)


# ---- metrics: get_batch_statistics / ap_per_class -------------------------------------------------
@pytest.mark.parametrize("name", ["metrics_small", "metrics_mid"])
def test_metrics_oracle_against_reference_vectors(name):
    import numpy as np
    d = load(name)
    dets = unpack_list(d, "dets")
    tg = T(d["targets"])
    stats = rp.get_batch_statistics(dets, tg, float(d["thr"]))
    assert len(stats) == int(d["n_stats"])
    for i, st in enumerate(stats):
        assert st[0].dtype == np.float64 and np.array_equal(st[0], d[f"stat_tp_{i}"])
    tp, sc, lb = [np.concatenate(x, 0) for x in zip(*stats)]
    assert np.array_equal(tp, d["tp"]) and np.array_equal(sc, d["scores"]) and np.array_equal(lb, d["labels"])
    p, r, ap, f1, cls = rp.ap_per_class(tp, sc, lb, d["target_cls"].tolist())
    assert np.array_equal(cls, d["classes"]) and cls.dtype == np.int32
    for got, want in ((p, d["p"]), (r, d["r"]), (ap, d["ap"]), (f1, d["f1"])):
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=0)


# ---- get_yolo_statistics ---------------------------------------------------------------------------
def _stats_self(kind, d):
    import types
    C, img = int(d["C"]), int(d["img"])
    if kind == "v3":
        anchors = [[tuple(map(float, a)) for a in lvl] for lvl in d["v3_anchors"]]
        return types.SimpleNamespace(anch_masks=None, anchors=anchors, num_classes=C, img_size=img, ignore_thres=0.5)
    if kind == "v2":
        return types.SimpleNamespace(anch_masks=None, anchors=[tuple(map(float, a)) for a in d["v2_anchors"]], num_classes=C,
                                     img_size=img, ignore_thres=0.5)
    return types.SimpleNamespace(anch_masks=d["v4_masks"].tolist(), anchors=d["v4_anchors"], num_classes=C, img_size=img,
                                 ignore_thres=0.5)


def test_yolo_statistics_oracle_against_reference_vectors():
    import numpy as np
    d = load("yolo_stats")
    tg = T(d["target"])
    heads = [T(d[f"v3_head_{G}"]) for G in (13, 26, 52)]
    for kind, hs, grids in (("v3", heads, (13, 26, 52)), ("v4", heads[:2], (13, 26))):
        bm = rp.get_yolo_statistics(_stats_self(kind, d), [h.clone() for h in hs], tg)
        for G in grids:
            np.testing.assert_allclose([float(x) for x in bm[G][:6]], d[f"{kind}_metrics_{G}"], rtol=1e-6, atol=1e-7)
            assert torch.equal(bm[G][6], T(d[f"{kind}_output_{G}"]))
    bm = rp.get_yolo_statistics(_stats_self("v2", d), T(d["v2_head"]), tg)
    np.testing.assert_allclose([float(x) for x in bm[13][:6]], d["v2_metrics"], rtol=1e-6, atol=1e-7)
    assert torch.equal(bm[13][6], T(d["v2_output"]))


# ---- fused v5 loss terms ------------------------------------------------------------------------
@pytest.mark.parametrize("tag,nc", [("", 4), ("1c", 1)])
def test_v5_loss_oracle_against_reference_vectors(tag, nc):
    import numpy as np
    d = load("v5_loss")
    pk, gk, tk, mk = (("p_", "grad_", "target", "metrics") if not tag else ("p1c_", "grad1c_", "target_1c", "metrics_1c"))
    p = [T(d[f"{pk}{i}"]).clone().requires_grad_(True) for i in range(3)]
    m = rp.v5_loss(p, T(d[tk]), T(d["anchors_scaled"]), 3, 3, nc)
    m["loss"].backward()
    got = [float(m[k].detach()) for k in ("loss", "Localization", "Classification", "Conf_obj")]
    np.testing.assert_allclose(got, d[mk], rtol=1e-6, atol=1e-7)
    for i in range(3):
        torch.testing.assert_close(p[i].grad, T(d[f"{gk}{i}"]), rtol=1e-5, atol=1e-8)

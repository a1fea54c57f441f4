"""GPU parity: loss-side target assignment (accuracy.py:305-380, 472-521; losses.py:105-123, 199-218,
423-443) against the golden vectors of the unmodified reference and the CPU oracle."""
import pytest
import torch

import objectdetectionpl_b200 as od
from objectdetectionpl_b200 import synth
from oracle import ref_port as rp
from tests.golden_io import load, T

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BT_NAMES = ["iou_scores", "class_mask", "obj_mask", "noobj_mask", "tx", "ty", "tw", "th", "tcls", "tconf"]


@pytest.mark.parametrize("name", ["bt_g13", "bt_g26_dups"])
def test_build_targets_golden(name):
    d = load(name)
    out = od.build_targets(T(d["pred_boxes"]).to(DEV), T(d["pred_cls"]).to(DEV), T(d["target"]).to(DEV),
                           T(d["anchors"]).to(DEV), 0.5)
    for k, v in zip(BT_NAMES, out):
        w = T(d[k])
        assert v.dtype == w.dtype and tuple(v.shape) == tuple(w.shape), k
        if k in ("tw", "th"):   # logf vs torch.log: last ulp
            torch.testing.assert_close(v.cpu(), w, rtol=1e-5, atol=1e-6, msg=k)
        else:
            assert torch.equal(v.cpu(), w), k


def test_build_targets_oracle_large_and_guards():
    B, A, G, C = 8, 3, 52, 80
    g = torch.Generator().manual_seed(5)
    tg = synth.labels(B, C, 401, max_per_image=100)
    pb = torch.rand(B, A, G, G, 4, generator=g) * G
    pc = torch.rand(B, A, G, G, C, generator=g)
    an = torch.tensor([[1.25, 1.625], [2.0, 3.75], [4.125, 2.875]])
    want = rp.build_targets(pb, pc, tg, an, 0.5)
    got = od.build_targets(pb.to(DEV), pc.to(DEV), tg.to(DEV), an.to(DEV), 0.5)
    for k, v, w in zip(BT_NAMES, got, want):
        if k in ("tw", "th"):
            torch.testing.assert_close(v.cpu(), w, rtol=1e-5, atol=1e-6, msg=k)
        else:
            assert torch.equal(v.cpu(), w), k
    # index guard (accuracy.py:340-344): one out-of-range image id suppresses every scatter but the per-target
    # ignore-threshold clearing of the valid rows
    bad = tg.clone()
    bad[3, 0] = B + 2
    want = rp.build_targets(pb, pc, bad, an, 0.5)
    got = od.build_targets(pb.to(DEV), pc.to(DEV), bad.to(DEV), an.to(DEV), 0.5)
    for k, v, w in zip(BT_NAMES, got, want):
        assert torch.equal(v.cpu(), w), "guard/" + k
    # label guard (accuracy.py:365-367)
    bad = tg.clone()
    bad[7, 1] = C + 1
    want = rp.build_targets(pb, pc, bad, an, 0.5)
    got = od.build_targets(pb.to(DEV), pc.to(DEV), bad.to(DEV), an.to(DEV), 0.5)
    for k, v, w in zip(BT_NAMES, got, want):
        assert torch.equal(v.cpu(), w), "label-guard/" + k


@pytest.mark.parametrize("name", ["btv5_small", "btv5_mid"])
def test_build_targets_v5_golden(name):
    d = load(name)
    shapes = [tuple(d[f"p_{i}"].shape) if name == "btv5_small" else tuple(int(x) for x in d[f"p_{i}"]) for i in range(3)]
    tcls, tbox, idx, anch = od.build_targets_v5(shapes, T(d["target"]).to(DEV), T(d["anchors"]), 3, 3)
    for i in range(3):
        assert tcls[i].dtype == torch.int64 and idx[i][0].dtype == torch.int64
        assert torch.equal(tcls[i].cpu(), T(d[f"tcls_{i}"]))
        assert torch.equal(tbox[i].cpu(), T(d[f"tbox_{i}"]))
        assert torch.equal(anch[i].cpu(), T(d[f"anch_{i}"]))
        for k, nm in enumerate("b a gj gi".split()):
            assert torch.equal(idx[i][k].cpu(), T(d[f"{nm}_{i}"])), (i, nm)


@pytest.mark.parametrize("B", [64, 160])        # 3 x nt = 9.4 k pairs (one compaction round) / ~24 k (two rounds, ragged tail)
def test_build_targets_v5_full_batch_and_empty(B):
    C, img = 80, 640
    tg = synth.labels(B, C, 4, max_per_image=100)
    stride = torch.tensor([8., 16., 32.])
    anchors = torch.tensor(synth.YOLOV5_ANCHORS).float().view(3, -1, 2) / stride.view(-1, 1, 1)
    shapes = [(B, 3, img // s, img // s, 5 + C) for s in (8, 16, 32)]
    want = rp.build_targets_v5(shapes, tg, anchors, 3, 3)
    got = od.build_targets_v5(shapes, tg.to(DEV), anchors, 3, 3)
    for i in range(3):
        assert torch.equal(got[0][i].cpu(), want[0][i])
        assert torch.equal(got[1][i].cpu(), want[1][i])
        assert torch.equal(got[3][i].cpu(), want[3][i])
        for k in range(4):
            assert torch.equal(got[2][i][k].cpu(), want[2][i][k])
    empty = od.build_targets_v5(shapes, torch.zeros(0, 6, device=DEV), anchors, 3, 3)
    assert all(t.shape[0] == 0 for t in empty[0]) and all(t.shape == (0, 4) for t in empty[1])


def test_v5_match_level_golden_lbox_and_grad():
    d = load("btv5_small")
    tg, anchors = T(d["target"]), T(d["anchors"])
    p = [T(d[f"p_{i}"]).to(DEV).requires_grad_(True) for i in range(3)]
    tcls, tbox, idx, anch = od.build_targets_v5(p, tg.to(DEV), anchors, 3, 3)
    pc = [T(d[f"p_{i}"]) for i in range(3)]
    wt = rp.build_targets_v5([t.shape for t in pc], tg, anchors, 3, 3)
    lbox = 0
    for i in range(3):
        giou, tobj = od.v5_match_level(p[i], tbox[i], idx[i], anch[i])
        wg, wo = rp.v5_match_level(pc[i], wt[1][i], wt[2][i], wt[3][i])
        torch.testing.assert_close(giou.detach().cpu(), wg, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(tobj.cpu(), wo, rtol=1e-5, atol=1e-6)
        assert torch.equal(tobj.cpu() != 0, wo != 0)
        lbox = lbox + (1.0 - giou).mean()
    lbox = lbox * 0.05
    torch.testing.assert_close(lbox.detach().cpu().reshape(1), T(d["lbox"]), rtol=1e-5, atol=1e-7)
    lbox.backward()
    for i in range(3):
        torch.testing.assert_close(p[i].grad.cpu(), T(d[f"lbox_grad_{i}"]), rtol=2e-4, atol=1e-8)


def test_ssd_match_and_retina_assign_golden():
    d = load("match")
    idx, matched = od.ssd_match(T(d["priors"]).to(DEV), T(d["gt"]).to(DEV), 0.5)
    assert idx.dtype == torch.int64 and matched.dtype == torch.bool
    assert torch.equal(idx.cpu(), T(d["ssd_idx"]))
    assert torch.equal(matched.cpu(), T(d["ssd_matched"]))
    anchors, tg = T(d["r_anchors"]), T(d["r_target"])
    B, img = int(d["r_B"]), float(d["r_img"])
    loc, cls = od.retina_assign(anchors.to(DEV), tg.to(DEV), B, img)
    loc, cls = loc.cpu(), cls.cpu()
    assert cls.dtype == torch.int64
    assert torch.equal(cls[cls > -1], T(d["r_cls_nonignored"]))
    torch.testing.assert_close(loc[cls > 0], T(d["r_loc_pos"]), rtol=1e-5, atol=1e-6)


def test_retina_assign_oracle_unsorted_targets():
    anchors = synth.retina_priors(256)
    B = 4
    tg = synth.labels(B, 9, 77, max_per_image=20)
    tg[:, 4:6] = tg[:, 4:6] * 1.5 + 0.05
    perm = torch.randperm(tg.shape[0], generator=torch.Generator().manual_seed(1))
    tg = tg[perm]                              # image ids interleaved: order within an image must be preserved
    wl, wc = rp.retina_assign(anchors, tg, B, 256.0)
    gl, gc = od.retina_assign(anchors.to(DEV), tg.to(DEV), B, 256.0)
    assert torch.equal(gc.cpu(), wc)
    torch.testing.assert_close(gl.cpu(), wl, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("variant", ["small", "many", "degenerate", "nan", "inf", "odd_anchor"])
def test_retina_assign_hull_filter(variant):
    """The per-CTA hull filter must be invisible: small targets (most are filtered away), > 256 targets in one image
    (several chunks), exact duplicates (first maximum wins), and boxes that switch the filter off mid-image."""
    anchors = synth.retina_priors(256)
    B = 3
    n = 300 if variant == "many" else 40
    tg = synth.labels(B, 9, 91, max_per_image=n, min_per_image=n)
    tg[:, 4:6] = tg[:, 4:6] * 0.3 + 0.02
    tg[5] = tg[2]                                        # duplicate inside image 0
    k = n + n // 2                                       # a row in the middle of image 1
    if variant == "degenerate":
        tg[k, 4] = -0.5                                  # negative width -> negative area
        tg[k + 1, 4:6] = 0.0
    elif variant == "nan":
        tg[k, 2] = float("nan")
    elif variant == "inf":
        tg[k, 4] = float("inf")
    elif variant == "odd_anchor":
        anchors = anchors.clone()
        anchors[1000, 2] = float("inf")
        anchors[5000, 3] = -30.0
    wl, wc = rp.retina_assign(anchors, tg, B, 256.0)
    gl, gc = od.retina_assign(anchors.to(DEV), tg.to(DEV), B, 256.0)
    assert torch.equal(gc.cpu(), wc)
    torch.testing.assert_close(gl.cpu(), wl, rtol=1e-5, atol=1e-6, equal_nan=True)


def test_ssd_match_oracle_many_gt():
    pri = synth.ssd_priors()
    g = torch.Generator().manual_seed(9)
    gt = torch.cat([0.1 + torch.rand(40, 2, generator=g) * 0.8, 0.03 + torch.rand(40, 2, generator=g) * 0.5], 1)
    gt[7] = gt[3]                              # duplicate GT -> "last GT wins" on its forced prior
    wi, wm = rp.ssd_match(pri, gt, 0.5)
    gi, gm = od.ssd_match(pri.to(DEV), gt.to(DEV), 0.5)
    assert torch.equal(gi.cpu(), wi) and torch.equal(gm.cpu(), wm)


# ---- fused v5 loss terms (losses.py:98-152): values and gradients within 1e-5 relative -----------------
@pytest.mark.parametrize("tag,nc", [("", 4), ("1c", 1)])
def test_v5_loss_golden_reference_vectors(tag, nc):
    import numpy as np
    from tests.golden_io import load, T
    d = load("v5_loss")
    pk, gk, tk, mk = (("p_", "grad_", "target", "metrics") if not tag else ("p1c_", "grad1c_", "target_1c", "metrics_1c"))
    p = [T(d[f"{pk}{i}"]).to(DEV).requires_grad_(True) for i in range(3)]
    m = od.v5_loss(p, T(d[tk]).to(DEV), T(d["anchors_scaled"]).to(DEV), 3, 3, nc)
    assert all(m[k].shape == (1,) for k in m)
    m["loss"].backward()
    got = [float(m[k].detach()) for k in ("loss", "Localization", "Classification", "Conf_obj")]
    np.testing.assert_allclose(got, d[mk], rtol=1e-5, atol=1e-7)
    for i in range(3):
        torch.testing.assert_close(p[i].grad.cpu(), T(d[f"{gk}{i}"]), rtol=1e-4, atol=2e-8)


@pytest.mark.parametrize("B,C,img,seed", [(4, 20, 320, 1), (8, 80, 640, 2)])
def test_v5_loss_against_oracle_autograd(B, C, img, seed):
    from oracle import ref_port as rp
    g = torch.Generator().manual_seed(seed)
    p_cpu = [torch.randn(B, 3, img // s, img // s, 5 + C, generator=g) for s in (8, 16, 32)]
    tg = synth.labels(B, C, seed + 20, max_per_image=30)
    stride = torch.tensor([8., 16., 32.])
    anchors = torch.tensor(synth.YOLOV5_ANCHORS).float().view(3, -1, 2) / stride.view(-1, 1, 1)
    pw = [t.clone().requires_grad_(True) for t in p_cpu]
    want = rp.v5_loss(pw, tg, anchors, 3, 3, C)
    want["loss"].backward()
    pg = [t.to(DEV).requires_grad_(True) for t in p_cpu]
    got = od.v5_loss(pg, tg.to(DEV), anchors.to(DEV), 3, 3, C)
    got["loss"].backward()
    for k in want:
        torch.testing.assert_close(got[k].detach().cpu(), want[k].detach(), rtol=1e-5, atol=1e-7)
    for a, b in zip(pg, pw):
        torch.testing.assert_close(a.grad.cpu(), b.grad, rtol=1e-4, atol=2e-8)


def test_v5_loss_without_targets():
    from oracle import ref_port as rp
    p_cpu = [torch.randn(2, 3, 8, 8, 9), torch.randn(2, 3, 4, 4, 9)]
    anchors = torch.tensor([[[1.25, 1.6], [2.0, 3.75], [4.1, 2.9]], [[1.9, 3.8], [3.9, 2.8], [3.7, 7.4]]])
    tg = torch.zeros(0, 6)
    want = rp.v5_loss([t.clone() for t in p_cpu], tg, anchors, 2, 3, 4)
    got = od.v5_loss([t.to(DEV) for t in p_cpu], tg.to(DEV), anchors.to(DEV), 2, 3, 4)
    for k in want:
        torch.testing.assert_close(got[k].cpu(), want[k], rtol=1e-5, atol=1e-7)


def test_v5_loss_level_by_level_equals_fused_node():
    """`v5_loss_level` (one autograd node per level, reference-typed int64 indices) and `v5_loss` (one node for all levels)
    run the same kernels: same terms, and gradients through a mix of the four outputs agree."""
    B, C = 4, 12
    g = torch.Generator().manual_seed(31)
    p_cpu = [torch.randn(B, 3, 160 // s, 160 // s, 5 + C, generator=g) for s in (8, 16, 32)]
    tg = synth.labels(B, C, 5, max_per_image=12).to(DEV)
    stride = torch.tensor([8., 16., 32.])
    anchors = (torch.tensor(synth.YOLOV5_ANCHORS).float().view(3, -1, 2) / stride.view(-1, 1, 1)).to(DEV)
    pa = [t.to(DEV).requires_grad_(True) for t in p_cpu]
    pb = [t.to(DEV).requires_grad_(True) for t in p_cpu]
    fused = od.v5_loss(pa, tg, anchors, 3, 3, C)
    (fused["loss"] * 2.0 + fused["Localization"] * 3.0 - fused["Conf_obj"] + fused["Classification"] * 0.5).sum().backward()
    tcls, tbox, indices, anch = od.build_targets_v5(pb, tg, anchors, 3, 3)
    lbox = torch.zeros(1, device=DEV); lobj = torch.zeros(1, device=DEV); lcls = torch.zeros(1, device=DEV)
    for i in range(3):
        t_box, t_obj, t_cls, _ = od.v5_loss_level(pb[i], tbox[i], indices[i], anch[i], tcls[i])
        if indices[i][0].shape[0]:
            lbox = lbox + t_box
            lcls = lcls + t_cls
        lobj = lobj + t_obj
    lbox, lobj, lcls = lbox * 0.05, lobj * 1.0, lcls * 0.58
    loss = lbox + lobj + lcls
    (loss * 2.0 + lbox * 3.0 - lobj + lcls * 0.5).sum().backward()
    assert torch.equal(fused["loss"], loss) and torch.equal(fused["Localization"], lbox)
    assert torch.equal(fused["Classification"], lcls) and torch.equal(fused["Conf_obj"], lobj)
    for a, b in zip(pa, pb):
        torch.testing.assert_close(a.grad, b.grad, rtol=1e-6, atol=1e-12)


def test_build_targets_v5_device_anchors_cache_sees_updates():
    """Device-resident anchors are copied to the host once per tensor and version: an in-place change must be seen, and a
    different tensor that happens to reuse the storage must not hit the old entry."""
    tg = synth.labels(2, 5, 3, max_per_image=10).to(DEV)
    shapes = [(2, 3, 20, 20, 10)]
    a1 = torch.tensor([[[1.25, 1.6], [2.0, 3.75], [4.1, 2.9]]], device=DEV)
    want1 = od.build_targets_v5(shapes, tg, a1.cpu(), 1, 3)
    got1 = od.build_targets_v5(shapes, tg, a1, 1, 3)
    got1b = od.build_targets_v5(shapes, tg, a1, 1, 3)               # cached
    a1.mul_(3.0)                                                      # same object, new version
    want2 = od.build_targets_v5(shapes, tg, a1.cpu(), 1, 3)
    got2 = od.build_targets_v5(shapes, tg, a1, 1, 3)
    del a1
    a3 = torch.tensor([[[9.0, 9.0], [0.5, 0.5], [2.0, 2.0]]], device=DEV)   # may reuse a1's storage
    want3 = od.build_targets_v5(shapes, tg, a3.cpu(), 1, 3)
    got3 = od.build_targets_v5(shapes, tg, a3, 1, 3)
    for got, want in ((got1, want1), (got1b, want1), (got2, want2), (got3, want3)):
        assert torch.equal(got[0][0], want[0][0]) and torch.equal(got[1][0], want[1][0]) and torch.equal(got[3][0], want[3][0])
        assert all(torch.equal(x, y) for x, y in zip(got[2][0], want[2][0]))
    assert got1[0][0].shape != got2[0][0].shape or not torch.equal(got1[3][0], got2[3][0])

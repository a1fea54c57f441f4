"""GPU: the drop-in EXECUTED.  The unmodified reference (staged byte for byte into oracle/_ref by oracle/stage_ref.py,
or /root/reference in the build container) runs its own code on CUDA tensors twice — stock, and after
`objectdetectionpl_b200.install(...)` patched the seams INTEGRATION.md names — and the results are compared:

  * every model's `non_max_suppression`   model/YOLOV2.py:159, YOLOV3.py:273, YOLOV4.py:221, YOLOV5.py:157, SSD.py:249,
                                           RetinaNet.py:117
  * `MultiScaleRegionLoss_v5.forward`      LightningFunc/losses.py:98-152   (globals build_targets_v5 / bbox_iou_v5)
  * `RegionLoss_v3.forward`                LightningFunc/losses.py:636-736  (self.build_targets, copied at construction)
  * the body of `test_step`                LightningFunc/step.py:64-100     (self.non_max_suppression + get_batch_statistics,
                                           and the get_yolo_statistics branch)
"""
import types

import numpy as np
import pytest
import torch

import objectdetectionpl_b200 as od
from objectdetectionpl_b200 import synth
from oracle import ref_harness as rh
from oracle import stage_ref
from tests.golden_io import assert_boxes_close

DEV = "cuda:0"


def test_staged_reference_is_the_unmodified_reference():
    """CPU-runnable: whatever tree the harness imports is byte-identical to the manifest taken from /root/reference."""
    if not rh.available():
        pytest.skip("no reference tree: run `python -c 'import __graft_entry__ as g; g.build()'` in the build container (stages oracle/_ref)")
    ok, bad = stage_ref.verify(rh.REF_ROOT)
    assert ok, bad


@pytest.fixture(scope="module")
def ref():
    if not rh.available():
        pytest.skip("no reference tree on this box (oracle/_ref is staged by build() where /root/reference exists)")
    mods = dict(losses=rh.losses(), acc=rh.accuracy(), step=rh.ref_import("LightningFunc.step"))
    for v in (2, 3, 4, 5):
        mods[f"v{v}"] = getattr(rh.ref_import(f"model.YOLOV{v}"), f"YOLOv{v}")
    mods["ssd"] = rh.ref_import("model.SSD").SSD
    mods["retina"] = rh.ref_import("model.RetinaNet").RetinaNet
    return types.SimpleNamespace(**mods)


class _Installed:
    """install() on the reference's classes / modules for the duration of a `with` block, then restore the stock seams."""

    def __init__(self, ref):
        self.ref = ref

    def __enter__(self):
        import sys
        r = self.ref
        self.classes = [r.v2, r.v3, r.v4, r.v5, r.ssd, r.retina]
        self.saved_cls = [(c, c.__dict__.get("non_max_suppression"), c.__dict__.get("get_yolo_statistics")) for c in self.classes]
        self.saved_glob = [(sys.modules[c.__module__], sys.modules[c.__module__].get_yolo_statistics) for c in (r.v2, r.v3, r.v4)]
        names = ("build_targets_v5", "bbox_iou_v5", "build_targets", "bbox_iou", "iou", "xywh2xyxy", "get_batch_statistics",
                 "ap_per_class")
        self.saved_mod = [(m, {k: getattr(m, k) for k in names if hasattr(m, k)}) for m in (r.losses, r.acc, r.step)]
        od.install(*self.classes, losses_module=r.losses, accuracy_module=r.acc, step_module=r.step)
        return self

    def __exit__(self, *a):
        for c, nms, ys in self.saved_cls:
            c.non_max_suppression = nms
            if ys is not None:
                c.get_yolo_statistics = ys
            elif "get_yolo_statistics" in c.__dict__:
                delattr(c, "get_yolo_statistics")
        for m, f in self.saved_glob:
            m.get_yolo_statistics = f
        for m, d in self.saved_mod:
            for k, v in d.items():
                setattr(m, k, v)


def _compare_yolo_lists(got, want, what):
    assert len(got) == len(want), what
    for b, (g, w) in enumerate(zip(got, want)):
        assert (g is None) == (w is None), f"{what}[{b}]"
        if w is None:
            continue
        assert g.is_cuda and tuple(g.shape) == tuple(w.shape), f"{what}[{b}]: {tuple(g.shape)} vs {tuple(w.shape)}"
        assert torch.equal(g[:, 4:].cpu(), w[:, 4:].cpu()), f"{what}[{b}]: conf / class columns differ"
        assert_boxes_close(g[:, :4], w[:, :4], what=f"{what}[{b}]")


@pytest.mark.gpu
@pytest.mark.parametrize("ver,A,C,grids,img", [(5, 3, 6, [20, 10, 5], 160), (3, 3, 5, [5, 10, 20], 160), (4, 3, 4, [6, 12], 96),
                                               (2, 5, 7, [13], 416)])
def test_yolo_nms_stock_vs_installed_on_cuda(ref, ver, A, C, grids, img):
    cls = getattr(ref, f"v{ver}")
    lv = [t.to(DEV) for t in synth.yolo_planar(2, A, C, grids, img, 400 + ver, v5_view=(ver == 5))]
    arg = lambda: [t.clone() for t in lv] if len(lv) > 1 else lv[0].clone()
    want = cls.non_max_suppression(None, arg())                          # the reference's own loop, on CUDA tensors
    with _Installed(ref):
        assert cls.non_max_suppression in (od.non_max_suppression, od.non_max_suppression_v2)
        got = cls.non_max_suppression(None, arg())
    assert cls.non_max_suppression not in (od.non_max_suppression, od.non_max_suppression_v2)
    _compare_yolo_lists(got, want, f"YOLOv{ver}")


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["ssd", "retina"])
def test_prior_nms_stock_vs_installed_on_cuda(ref, which):
    cls = getattr(ref, which)
    pri = (synth.ssd_priors() if which == "ssd" else synth.retina_priors(128)).to(DEV)
    loc, c = synth.prior_heads(3, pri.shape[0], 5, 77, cls_mean=-3.0)
    me = types.SimpleNamespace(iou_boxes=pri)
    want = cls.non_max_suppression(me, (loc.to(DEV), c.to(DEV)))
    with _Installed(ref):
        got = cls.non_max_suppression(me, (loc.to(DEV), c.to(DEV)))
    assert len(got) == len(want)
    for b, (g, w) in enumerate(zip(got, want)):
        assert tuple(g.shape) == tuple(w.shape), f"{which}[{b}]"
        assert torch.equal(g[:, 6].cpu(), w[:, 6].cpu()) and torch.equal(g[:, 4].cpu(), w[:, 4].cpu())
        torch.testing.assert_close(g[:, 5].cpu(), w[:, 5].cpu(), rtol=1e-6, atol=1e-7)      # sigmoid on two code paths
        torch.testing.assert_close(g[:, :4].cpu(), w[:, :4].cpu(), rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_v5_criterion_forward_backward_stock_vs_installed(ref):
    """MultiScaleRegionLoss_v5.forward resolves build_targets_v5 / bbox_iou_v5 as module globals at call time
    (losses.py:102,118): the same criterion object runs stock, then with the globals patched."""
    B, C, img = 4, 6, 160
    crit = ref.losses.MultiScaleRegionLoss_v5(synth.YOLOV5_ANCHORS, None, None, None, None, C, img)
    g = torch.Generator().manual_seed(5)
    heads = [torch.randn(B, 3, img // s, img // s, 5 + C, generator=g) for s in (8, 16, 32)]
    # two well separated boxes per image: build_targets_v5 then yields no duplicate (image, anchor, cell) rows, so the
    # reference's own `tobj[b, a, gj, gi] = ...` scatter (losses.py:123) has one writer per cell.  (With duplicates the winner of
    # torch's CUDA index_put_ is not defined — the stock criterion on the CPU and the same criterion on CUDA then differ in
    # `Conf_obj` by ~1e-4 whatever build_targets_v5 is used.)
    tg = torch.tensor([[b, (b + k) % C, 0.27 + 0.46 * k, 0.31 + 0.4 * ((b + k) % 2), 0.11 + 0.02 * b, 0.17 + 0.03 * k]
                       for b in range(B) for k in range(2)], dtype=torch.float32)
    chk = od.build_targets_v5([(B, 3, img // s, img // s, 5 + C) for s in (8, 16, 32)], tg.to(DEV), crit.anchors, 3, 3)[2]
    rows = 0
    for bb, aa, gj, gi in chk:
        cell = ((bb * 3 + aa) * 64 + gj) * 64 + gi
        assert cell.unique().numel() == cell.numel(), "the test labels must not produce duplicate cells"
        rows += cell.numel()
    assert rows >= 20

    def run(c, dev):
        p = [h.detach().clone().to(dev).requires_grad_(True) for h in heads]
        m = c(p, tg.clone().to(dev))
        m["loss"].sum().backward()
        return {k: v.detach().float().cpu() for k, v in m.items()}, [t.grad.cpu() for t in p]

    try:
        want_m, want_g = run(crit, DEV)
        stock_on = "cuda"
    except RuntimeError as e:
        # torch >= 2.x: the reference's own build_targets_v5 indexes a CUDA tensor with a CPU index tensor it builds at
        # accuracy.py:477 ("indices should be either on cpu or on the same device"), i.e. the STOCK criterion cannot run on
        # CUDA tensors at all with this torch.  The stock half then runs the same unmodified code on the CPU.
        assert "indices should be" in str(e), e
        with rh.cpu_only():
            want_m, want_g = run(ref.losses.MultiScaleRegionLoss_v5(synth.YOLOV5_ANCHORS, None, None, None, None, C, img), "cpu")
        stock_on = "cpu"
    with _Installed(ref):
        assert ref.losses.build_targets_v5 is od.build_targets_v5 and ref.losses.bbox_iou_v5 is od.bbox_iou_v5
        got_m, got_g = run(crit, DEV)
    print(f"stock criterion ran on {stock_on}")
    assert set(got_m) == set(want_m)
    for k in want_m:
        torch.testing.assert_close(got_m[k], want_m[k], rtol=1e-5, atol=1e-7, msg=lambda s, k=k: f"{k}: {s}")
    for a, b in zip(got_g, want_g):
        # gradients of a mean over ~1e5 cells are ~1e-7 per element: the absolute floor is relative to the largest one
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5 * float(b.abs().max()))


@pytest.mark.gpu
def test_region_loss_v3_forward_stock_vs_installed(ref):
    """RegionLoss_v3 copies `build_targets` into self at construction (losses.py:654): one criterion is built before and
    one after install(); both run the reference's own forward (losses.py:668-736) on the same CUDA head."""
    B, C, G, img = 4, 5, 13, 416
    anchors = [(116, 90), (156, 198), (373, 326)]
    head = synth.raw_logits(B, 3, C, G, 8).to(DEV)
    tg = synth.labels(B, C, 9, max_per_image=6).to(DEV)
    mk = lambda: ref.losses.RegionLoss_v3(anchors, torch.nn.BCELoss, torch.nn.MSELoss, torch.nn.BCELoss, C, img)

    def run(crit):
        x = head.clone().requires_grad_(True)
        out = crit(x, tg.clone())
        out[0].backward()
        return [o.detach().float().cpu() for o in out], x.grad.cpu()

    want, want_g = run(mk())
    with _Installed(ref):
        crit = mk()
        assert crit.build_targets is od.build_targets
        got, got_g = run(crit)
    for i, (a, b) in enumerate(zip(got, want)):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-7, msg=lambda s, i=i: f"output {i}: {s}")
    torch.testing.assert_close(got_g, want_g, rtol=1e-4, atol=1e-5 * float(want_g.abs().max()))


def _fake_module(ref, checkname, heads, nms, img):
    """Just enough of a LightningModule for the unmodified `test_step` body (step.py:64-100)."""
    me = types.SimpleNamespace(checkname=checkname, img_size=img, inference=False)
    me.forward = lambda x: [h.clone() for h in heads]
    me.non_max_suppression = lambda out: nms(me, out)
    me.mark_target = lambda im, y, idx: im
    me.mark_pred = lambda im, dets: im
    return me


@pytest.mark.gpu
def test_test_step_body_stock_vs_installed_yolov5_branch(ref):
    """step.py:64-95: forward -> self.non_max_suppression -> get_batch_statistics (the RetinaNet / SSD / YOLOv5 branch)."""
    B, C, img = 2, 4, 64
    heads = [t.to(DEV) for t in synth.yolo_planar(B, 3, C, [8, 4, 2], img, 21, v5_view=True)]
    x = torch.rand(B, 3, img, img)
    g = torch.Generator().manual_seed(3)
    # labels near some of the stock detections so that true positives exist; test_step multiplies y[:, 2:] by img_size
    # once PER IMAGE of the batch (step.py:81, inside the loop), so the labels are pre-divided by img_size ** B
    dets = ref.v5.non_max_suppression(None, [h.clone() for h in heads])
    y = []
    for b, d in enumerate(dets):
        pick = d[torch.randperm(d.shape[0], generator=g)[:6].to(d.device)].cpu()
        y.append(torch.cat([torch.full((6, 1), float(b)), pick[:, 6:7], (pick[:, :4] + torch.randn(6, 4, generator=g)) / img ** B], 1))
    y = torch.cat(y)

    def run(nms):
        out = ref.step.test_step(_fake_module(ref, "YOLOv5", heads, nms, img), (x.clone(), y.clone()), 0)
        return out["sample_metric"], out["label"]

    want, want_l = run(ref.v5.non_max_suppression)
    with _Installed(ref):
        assert ref.step.get_batch_statistics is od.get_batch_statistics
        got, got_l = run(ref.v5.non_max_suppression)
    assert got_l == want_l and len(got) == len(want)
    assert sum(float(np.asarray(b[0]).sum()) for b in want) >= 4, "the stock run found no true positives: test is vacuous"
    for a, b in zip(got, want):
        assert len(a) == 3
        np.testing.assert_array_equal(np.asarray(a[0]), np.asarray(b[0]))          # true-positive flags
        np.testing.assert_array_equal(np.asarray(a[1]), np.asarray(b[1]))          # confidences
        np.testing.assert_array_equal(np.asarray(a[2]), np.asarray(b[2]))          # labels


@pytest.mark.gpu
def test_test_step_body_stock_vs_installed_yolov3_branch(ref):
    """step.py:96-100: the YOLOv2-v4 branch calls self.get_yolo_statistics (accuracy.py:382-470), set on the class."""
    B, C, G, img = 2, 4, 13, 416
    anchors = [(116 / 32, 90 / 32), (156 / 32, 198 / 32), (373 / 32, 326 / 32)]        # model/YOLOV3.py:55-56
    head = synth.raw_logits(B, 3, C, G, 61).to(DEV)
    tg = synth.labels(B, C, 62, max_per_image=6)
    tg[:, 2:] /= float(img) ** B            # test_step scales y[:, 2:] by img_size once per image (step.py:81) before the call
    x = torch.rand(B, 3, 32, 32)

    import sys
    v3mod = sys.modules[ref.v3.__module__]

    def run(cls):
        me = _fake_module(ref, "YOLOv3", [head], lambda s, out: cls.non_max_suppression(None, out), img)
        me.anch_masks, me.anchors, me.num_classes, me.ignore_thres = None, [anchors, anchors, anchors], C, 0.5
        # what the model's constructor does (model/YOLOV3.py:252): bind the module's global onto the object
        fn = v3mod.get_yolo_statistics
        me.get_yolo_statistics = lambda out, y: fn(me, out, y.to(DEV))
        return ref.step.test_step(me, (x.clone(), tg.clone()), 0)["sample_metric"]

    want = run(ref.v3)
    with _Installed(ref):
        assert ref.v3.get_yolo_statistics is od.get_yolo_statistics and v3mod.get_yolo_statistics is od.get_yolo_statistics
        got = run(ref.v3)
    assert v3mod.get_yolo_statistics is not od.get_yolo_statistics
    assert set(got) == set(want) == {G}
    assert all(np.isfinite(float(b)) for b in want[G][:6]), "stock metrics are not finite: test is vacuous"
    for a, b in zip(got[G][:6], want[G][:6]):
        np.testing.assert_allclose(float(a), float(b), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(got[G][6].cpu(), want[G][6].cpu(), rtol=1e-5, atol=1e-6)

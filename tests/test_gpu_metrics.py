"""GPU parity of the test-time metrics through the C ABI:
  get_batch_statistics — true-positive flags bit-exact against the reference's golden vectors and the oracle;
  ap_per_class         — precision / recall / AP / F1 in fp64 within 1e-12 relative (summation order differs)."""
import numpy as np
import pytest
import torch

import objectdetectionpl_b200 as od
from objectdetectionpl_b200 import synth
from oracle import ref_port as rp
from tests.golden_io import load, unpack_list, T

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("name", ["metrics_small", "metrics_mid"])
def test_golden_reference_vectors(name):
    d = load(name)
    dets = [None if t is None else t.to(DEV) for t in unpack_list(d, "dets")]
    stats = od.get_batch_statistics(dets, T(d["targets"]).to(DEV), float(d["thr"]))
    assert len(stats) == int(d["n_stats"])
    for i, st in enumerate(stats):
        assert st[0].dtype == np.float64
        assert np.array_equal(st[0], d[f"stat_tp_{i}"]), f"image {i}: true positives differ"
    tp, sc, lb = [np.concatenate(x, 0) for x in zip(*stats)]
    assert np.array_equal(sc, d["scores"]) and np.array_equal(lb, d["labels"])
    p, r, ap, f1, cls = od.ap_per_class(tp, sc, lb, d["target_cls"].tolist())
    assert cls.dtype == np.int32 and np.array_equal(cls, d["classes"])
    for got, want, nm in ((p, d["p"], "p"), (r, d["r"], "r"), (ap, d["ap"], "ap"), (f1, d["f1"], "f1")):
        assert got.dtype == np.float64
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-300, err_msg=nm)


def _random_case(B, C, K, M, seed, img=320.0):
    g = torch.Generator().manual_seed(seed)
    dets, tg = [], []
    for b in range(B):
        k = int(torch.randint(max(1, K // 2), K + 1, (1,), generator=g))
        xy = torch.rand(k, 2, generator=g) * img * 0.8
        wh = 8 + torch.rand(k, 2, generator=g) * img * 0.2
        conf = torch.rand(k, generator=g).sort(descending=True).values
        rows = torch.cat([xy, xy + wh, conf[:, None], torch.rand(k, 1, generator=g),
                          torch.randint(0, C, (k, 1), generator=g).float()], 1)
        dets.append(rows)
        m = int(torch.randint(0, M + 1, (1,), generator=g))
        if m:
            src = rows[torch.randint(0, k, (m,), generator=g)]
            box = src[:, :4] + torch.randn(m, 4, generator=g) * 4
            lab = torch.where(torch.rand(m, generator=g) < 0.8, src[:, 6], torch.randint(0, C, (m,), generator=g).float())
            tg.append(torch.cat([torch.full((m, 1), float(b)), lab[:, None], box], 1))
    tg = torch.cat(tg) if tg else torch.zeros(0, 6)
    tg = tg[torch.randperm(tg.shape[0], generator=g)]            # targets need not be grouped by image
    return dets, tg


@pytest.mark.parametrize("B,C,K,M,seed", [(4, 3, 60, 10, 1), (8, 20, 400, 40, 2), (2, 1, 700, 300, 3), (3, 5, 50, 0, 4)])
def test_batch_statistics_against_oracle(B, C, K, M, seed):
    dets, tg = _random_case(B, C, K, M, seed)
    if seed == 1:
        dets[2] = None
    want = rp.get_batch_statistics(dets, tg, 0.5)
    got = od.get_batch_statistics([None if t is None else t.to(DEV) for t in dets], tg.to(DEV), 0.5)
    assert len(got) == len(want)
    for i, (a, b) in enumerate(zip(got, want)):
        assert np.array_equal(a[0], b[0]), f"entry {i}: tp"
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


def test_batch_statistics_raw_on_padded_nms_output():
    """The device-resident form consumes the padded [B, n_pad, 7] rows + counts of the NMS directly (no host sync)."""
    levels = synth.yolo_planar(B=3, A=3, C=4, grids=[10, 5], img=80, seed=9, v5_view=True)
    rows, _, count = od.yolo_nms_raw([t.to(DEV) for t in levels], 3)
    dets = od.non_max_suppression(None, [t.to(DEV) for t in levels])
    tg = torch.cat([torch.cat([torch.full((5, 1), float(b)), d[:5, 6:7].cpu(), d[:5, :4].cpu() + 1.5], 1)
                    for b, d in enumerate(dets)])
    n_pad = rows.shape[1]
    row_start = torch.arange(3, device=DEV, dtype=torch.int64) * n_pad
    tp = od.batch_statistics_raw(rows, row_start, count, n_pad, tg.to(DEV), 0.5)
    want = rp.get_batch_statistics([d.cpu() for d in dets], tg, 0.5)
    for b in range(3):
        k = int(count[b])
        assert np.array_equal(tp[b, :k].double().cpu().numpy(), want[b][0])


# > 53 248 detections and <= 255 evaluated classes take the class-partitioned route (254 + the absent class = 255 is its
# widest case), 300 classes the streaming one on the same sort; the last has more detections than a headline batch keeps
@pytest.mark.parametrize("n,C,seed", [(1, 1, 1), (300, 3, 2), (5000, 20, 3), (70000, 80, 4), (1500, 300, 5),
                                      (60000, 254, 7), (60000, 300, 8), (1300000, 80, 6)])
def test_ap_per_class_against_oracle(n, C, seed):
    g = torch.Generator().manual_seed(seed)
    conf = ((torch.randperm(n, generator=g).double() + 0.5) / n).float()     # distinct values: numpy's argsort is unstable on ties
    assert torch.unique(conf).numel() == n
    cls = torch.randint(0, C + 1, (n,), generator=g).float()     # class C never appears among the targets
    tp = (torch.rand(n, generator=g) < 0.3).float()
    target_cls = torch.randint(0, C, (max(3, n // 4),), generator=g).float().tolist() + [float(C + 5)]
    want = rp.ap_per_class(tp.numpy().astype(np.float64), conf.numpy(), cls.numpy(), target_cls)
    got = od.ap_per_class(tp.numpy().astype(np.float64), conf.numpy(), cls.numpy(), target_cls)
    assert np.array_equal(got[4], want[4])
    for a, b in zip(got[:4], want[:4]):
        np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-300)
    assert (got[2] >= 0).all()        # (random tp flags may exceed the label count, so recall / AP can pass 1 here)


def test_ap_per_class_device_repeated_class():
    """A class listed twice (not what np.unique produces, but legal at the C ABI) gets the same numbers twice."""
    n = 80000
    g = torch.Generator().manual_seed(11)
    conf = ((torch.randperm(n, generator=g).double() + 0.5) / n).float().to(DEV)
    cls = torch.randint(0, 6, (n,), generator=g).float().to(DEV)
    tp = (torch.rand(n, generator=g) < 0.4).float().to(DEV)
    classes = torch.tensor([0, 3, 5, 3, 9], dtype=torch.int32, device=DEV)
    n_gt = torch.tensor([50, 70, 20, 70, 4], dtype=torch.int32, device=DEV)
    p, r, ap, f1 = (x.cpu() for x in od.ap_per_class_device(tp, conf, cls, classes, n_gt))
    assert p[1] == p[3] and r[1] == r[3] and ap[1] == ap[3] and f1[1] == f1[3] and ap[1] > 0
    assert p[4] == 0 and ap[4] == 0                                        # class 9 has no detections
    uniq = torch.tensor([0, 3, 5], dtype=torch.int32, device=DEV)
    p2, r2, ap2, f12 = (x.cpu() for x in od.ap_per_class_device(tp, conf, cls, uniq, n_gt[:3].contiguous()))
    assert torch.equal(p[:3], p2) and torch.equal(ap[:3], ap2) and torch.equal(r[:3], r2)


def test_ap_per_class_perfect_and_empty():
    tp = np.ones(10)
    conf = np.linspace(0.9, 0.1, 10).astype(np.float32)
    cls = np.zeros(10, np.float32)
    p, r, ap, f1, c = od.ap_per_class(tp, conf, cls, [0.0] * 10 + [3.0])
    assert c.tolist() == [0, 3]
    np.testing.assert_allclose([p[0], r[0], ap[0]], [1.0, 1.0, 1.0], rtol=1e-12)
    assert p[1] == 0 and r[1] == 0 and ap[1] == 0 and f1[1] == 0           # class without predictions (accuracy.py:238)


# ---- get_yolo_statistics: decoded map within 1e-5 relative (sigmoid / exp), metrics within 1e-5 ------
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


def _check_stats(bm, G, want_metrics, want_output):
    got = [float(x) for x in bm[G][:6]]
    np.testing.assert_allclose(got, want_metrics, rtol=1e-5, atol=1e-6)
    assert not bm[G][6].is_cuda and bm[G][6].shape == want_output.shape
    torch.testing.assert_close(bm[G][6], want_output, rtol=1e-5, atol=1e-5)
    assert all(isinstance(x, np.ndarray) and x.shape == () for x in bm[G][:6])


def test_yolo_statistics_golden_reference_vectors():
    d = load("yolo_stats")
    tg = T(d["target"]).to(DEV)
    heads = [T(d[f"v3_head_{G}"]).to(DEV) for G in (13, 26, 52)]
    for kind, hs, grids in (("v3", heads, (13, 26, 52)), ("v4", heads[:2], (13, 26))):
        s = _stats_self(kind, d)
        bm = od.get_yolo_statistics(s, [h.clone() for h in hs], tg)
        assert sorted(bm) == sorted(grids)
        for G in grids:
            _check_stats(bm, G, d[f"{kind}_metrics_{G}"], T(d[f"{kind}_output_{G}"]))
        assert s.num_anchors == 3 and s.grid_size == grids[-1]
    bm = od.get_yolo_statistics(_stats_self("v2", d), T(d["v2_head"]).to(DEV), tg)      # single tensor (accuracy.py:386)
    _check_stats(bm, 13, d["v2_metrics"], T(d["v2_output"]))


@pytest.mark.parametrize("B,C,G,seed", [(4, 20, 13, 1), (8, 80, 26, 2), (2, 3, 52, 3)])
def test_yolo_statistics_against_oracle(B, C, G, seed):
    import types
    head = synth.raw_logits(B, 3, C, G, seed)
    head.view(B, 3, 5 + C, G, G)[:, :, 4] += 3.0
    tg = synth.labels(B, C, seed + 10, max_per_image=20)
    anchors = [[(116, 90), (156, 198), (373, 326)]] * 3
    mk = lambda: types.SimpleNamespace(anch_masks=None, anchors=anchors, num_classes=C, img_size=416, ignore_thres=0.5)
    want = rp.get_yolo_statistics(mk(), [head.clone()], tg)
    got = od.get_yolo_statistics(mk(), [head.to(DEV)], tg.to(DEV))
    _check_stats(got, G, [float(x) for x in want[G][:6]], want[G][6])


def test_yolo_statistics_no_targets_gives_nan_means():
    import types
    head = synth.raw_logits(2, 3, 4, 13, 5).to(DEV)
    s = types.SimpleNamespace(anch_masks=None, anchors=[[(116, 90), (156, 198), (373, 326)]] * 3, num_classes=4, img_size=416,
                              ignore_thres=0.5)
    bm = od.get_yolo_statistics(s, [head], torch.zeros(0, 6, device=DEV))
    m = [float(x) for x in bm[13][:6]]
    assert np.isnan(m[0]) and np.isnan(m[4]) and m[1] == 0 and m[2] == 0 and m[3] == 0      # mean over an empty selection

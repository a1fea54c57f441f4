"""GPU parity on adversarial inputs: duplicates and huge clusters, NaN scores, negative thresholds, one-candidate
images, ragged level sizes — everything checked against the CPU oracle (which follows the reference loop)."""
import pytest
import torch

import objectdetectionpl_b200 as od
from objectdetectionpl_b200 import synth
from oracle import ref_port as rp
from tests.golden_io import assert_rows_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run(levels, A=3, **kw):
    got, gidx = od.non_max_suppression(None, [t.to(DEV) for t in levels], return_index=True, **kw)
    thr = kw["conf_thres"] if kw.get("compat", True) is False else -0.0151
    want, widx = rp.yolo_nms_rows(rp.yolo_rows_from_planar(levels, A), conf_thres=thr, nms_thres=kw.get("nms_thres", 0.4),
                                  return_index=True)
    for b in range(len(want)):
        assert (got[b] is None) == (want[b] is None)
        if want[b] is None:
            continue
        assert torch.equal(gidx[b].cpu(), widx[b]), f"image {b}"
        assert_rows_close(got[b], want[b], rtol=1e-5, atol=1e-4, what=f"image {b}")
    return got


def _planar(B, A, C, G):
    return torch.zeros(B, A, 5 + C, G, G)


def test_one_huge_cluster_of_identical_boxes():
    """Every candidate is the same box of the same class: one keeper, its box the conf-weighted mean of 768 rows
    (two 384-row chunks -> the cross-chunk merge path)."""
    t = _planar(1, 3, 2, 16)
    t[:, :, 0:2] = 50.0
    t[:, :, 2:4] = 20.0
    t[:, :, 4] = torch.linspace(0.2, 0.9, 3 * 256).view(1, 3, 16, 16)
    t[:, :, 5] = 0.7
    got = _run([t.view(1, 21, 16, 16)])
    assert got[0].shape[0] == 1


def test_nested_and_chained_boxes_single_class():
    """Concentric boxes shrinking by 5% per step: a long suppression chain inside one class segment."""
    n = 3 * 8 * 8
    t = _planar(1, 3, 1, 8)
    scale = 0.95 ** torch.arange(n, dtype=torch.float32)
    t[:, :, 0] = 100.0
    t[:, :, 1] = 100.0
    t[:, :, 2] = (150.0 * scale).view(1, 3, 8, 8)
    t[:, :, 3] = (150.0 * scale).view(1, 3, 8, 8)
    t[:, :, 4] = torch.linspace(0.9, 0.1, n).view(1, 3, 8, 8)
    t[:, :, 5] = 1.0
    _run([t.view(1, 18, 8, 8)])
    _run([t.view(1, 18, 8, 8)], nms_thres=0.9)
    _run([t.view(1, 18, 8, 8)], nms_thres=0.0)


def test_negative_nms_threshold_uses_the_exact_path():
    """nms_thres < 0: even non-overlapping same-class boxes suppress each other (IoU 0 > thr); the half2 pre-filter must
    not be used then."""
    lv = synth.yolo_planar(2, 3, 3, [8, 4], 64, 17)
    _run(lv, nms_thres=-0.5)


def test_nan_class_scores_and_nan_confidence():
    """torch.max returns the first NaN as the maximum; a NaN score sorts last; a NaN confidence fails `>= thr` and is
    dropped (model/YOLOV3.py:310-317)."""
    lv = synth.yolo_planar(1, 3, 4, [8], 64, 23)
    p = lv[0].view(1, 3, 9, 8, 8)
    p[0, 0, 6, 2, 3] = float("nan")         # class 1 of one candidate
    p[0, 1, 5, 4, 4] = float("nan")
    p[0, 1, 7, 4, 4] = float("nan")         # two NaN classes: the first wins
    p[0, 2, 4, 1, 1] = float("nan")         # NaN confidence -> filtered
    got, gidx = od.non_max_suppression(None, [t.to(DEV) for t in lv], return_index=True)
    rows = rp.yolo_rows_from_planar(lv, 3)
    # the reference's own loop would mis-handle NaN boxes, so check the documented pieces instead of the full loop
    g, gi = got[0].cpu(), gidx[0].cpu()
    dropped = 2 * 64 + 1 * 8 + 1
    assert dropped not in gi.tolist()
    n1 = 0 * 64 + 2 * 8 + 3
    n2 = 1 * 64 + 4 * 8 + 4
    for cand, cls in ((n1, 1.0), (n2, 0.0)):
        pos = (gi == cand).nonzero()
        assert pos.numel() == 1
        row = g[pos.item()]
        assert row[6].item() == cls and torch.isnan(row[5])
    assert set(gi[-2:].tolist()) == {n1, n2}                  # NaN scores sort last
    cls_conf, cls_id = rows[0, :, 5:].max(1)
    ok = torch.ones(rows.shape[1], dtype=torch.bool)
    ok[[dropped, n1, n2]] = False
    pos = {c: i for i, c in enumerate(gi.tolist())}
    for cand in ok.nonzero().flatten().tolist()[:50]:
        if cand in pos:
            assert g[pos[cand], 6].item() == float(cls_id[cand]) and g[pos[cand], 5].item() == float(cls_conf[cand])


def test_single_candidate_and_single_survivor():
    t = _planar(2, 1, 3, 1)                 # A=1, G=1: one candidate per image
    t[:, 0, 0:4, 0, 0] = torch.tensor([10.0, 10.0, 4.0, 4.0])
    t[:, 0, 4, 0, 0] = torch.tensor([0.5, 0.05])
    t[:, 0, 5:, 0, 0] = torch.tensor([0.1, 0.9, 0.3])
    from objectdetectionpl_b200.postprocess import _yolo_nms
    got = _yolo_nms([t.view(2, 8, 1, 1).to(DEV)], 1, 0.2, 0.4, False, None, None, None, False)
    assert got[0].shape == (1, 7) and got[1] is None
    assert got[0][0].tolist() == [8.0, 8.0, 12.0, 12.0, 0.5, pytest.approx(0.9), 1.0]


def test_ragged_levels_not_multiple_of_tile():
    """N = 3*(7*7 + 5*5 + 3*3) = 249 and N = 3*(20*20+13*13) = 1707: partial tiles, odd grids (scalar loads)."""
    _run(synth.yolo_planar(2, 3, 5, [7, 5, 3], 56, 31))
    _run(synth.yolo_planar(2, 3, 5, [20, 13], 160, 32))


def test_many_classes_two_radix_passes_and_limits():
    _run(synth.yolo_planar(1, 3, 700, [6, 3], 48, 33))
    with pytest.raises(ValueError):
        od.non_max_suppression(None, [torch.zeros(1, 3 * (5 + 4096), 2, 2, device=DEV)])


def test_output_views_do_not_alias_the_input_and_input_is_untouched():
    lv = [t.to(DEV) for t in synth.yolo_planar(2, 3, 4, [8, 4], 64, 35)]
    before = [t.clone() for t in lv]
    od.non_max_suppression(None, lv)
    for a, b in zip(lv, before):
        assert torch.equal(a, b)


@pytest.mark.parametrize("layout", ["planar", "channels_last"])
def test_arbitrary_bit_patterns_terminate(layout):
    """Heads made of random bit patterns (NaN, inf, denormals, huge / negative sizes) must run to completion in both
    layouts and give the same rows in both (the reference itself loops forever on w <= -1)."""
    dev = torch.device(DEV)
    g = torch.Generator(device=dev).manual_seed(11)
    for it, (B, A, C, grids) in enumerate([(2, 3, 20, [20, 10, 5]), (2, 3, 4, [40, 20, 10]), (1, 3, 80, [40, 20])]):
        levels = []
        for G in grids:
            bits = torch.randint(-2**31, 2**31 - 1, (B, A, 5 + C, G, G), device=dev, generator=g, dtype=torch.int64).to(torch.int32)
            v = bits.view(torch.float32).clone()
            if it % 2:
                v[:, :, 4:] = torch.rand(B, A, 1 + C, G, G, device=dev, generator=g)
            levels.append(v)
        planar = [v.reshape(v.shape[0], -1, v.shape[3], v.shape[4]) for v in levels]
        rows_p, _, count_p = od.yolo_nms_raw(planar, A)
        if layout == "channels_last":
            cl = [v.permute(0, 1, 3, 4, 2).contiguous() for v in levels]
            rows, _, count = od.yolo_nms_raw(cl, A, layout="channels_last")
            torch.cuda.synchronize()
            assert torch.equal(count, count_p)
            for b, k in enumerate(count.tolist()):
                assert torch.equal(rows[b, :k].view(torch.int32), rows_p[b, :k].view(torch.int32))     # bit patterns (NaN-safe)
        torch.cuda.synchronize()
        assert int(count_p.min()) >= 0


@pytest.mark.parametrize("nms_thres", [0.0, 0.05, 0.2, 0.45, 0.7, 0.95])
@pytest.mark.parametrize("C,seed", [(1, 3), (4, 4), (20, 5)])
def test_threshold_sweep_dense_overlaps(nms_thres, C, seed):
    """The half2 pre-filter folds the threshold into its per-row summary: sweep it from 0 to 0.95 on boxes that overlap
    heavily (few classes, boxes a third of the image wide, nested small-in-large pairs) and on ordinary heads."""
    lv = synth.yolo_planar(B=2, A=3, C=C, grids=[12, 6], img=96, seed=seed, v5_view=False)
    g = torch.Generator().manual_seed(seed)
    for t in lv:                                         # widths / heights from tiny to image-sized: nested pairs
        v = t.view(2, 3, 5 + C, t.shape[-1], t.shape[-1])
        v[:, :, 2:4] = 2.0 + torch.rand(v[:, :, 2:4].shape, generator=g) ** 3 * 94.0
    _run(lv, nms_thres=nms_thres)

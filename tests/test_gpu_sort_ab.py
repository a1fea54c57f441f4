"""GPU: the one-launch cluster sort (clustersort.cu) against the multi-launch look-back sort (segsort.cu).
Both must give bit-identical pipelines: same kept rows, same candidate indices, same counts — for every cluster
size (1, 2, 4, 8 CTAs per image), for two class passes (C > 256), for sparse survivors and for the prior path."""
import os

import pytest
import torch

import objectdetectionpl_b200 as od
from objectdetectionpl_b200 import synth
from objectdetectionpl_b200.postprocess import yolo_nms_raw, prior_nms_raw

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class _sort_mode:
    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        self.old = os.environ.get("B200DET_SORT")
        if self.mode is None:
            os.environ.pop("B200DET_SORT", None)
        else:
            os.environ["B200DET_SORT"] = self.mode

    def __exit__(self, *a):
        if self.old is None:
            os.environ.pop("B200DET_SORT", None)
        else:
            os.environ["B200DET_SORT"] = self.old


def _run_yolo(levels, A, conf_thres, mode):
    with _sort_mode(mode):
        rows, index, count = yolo_nms_raw(levels, A, conf_thres, 0.4, want_index=True)
        torch.cuda.synchronize()
    c = count.cpu()
    rows = rows.cpu()
    index = index.cpu()
    return [(rows[b, :k].clone(), index[b, :k].clone()) for b, k in enumerate(c.tolist())]


@pytest.mark.parametrize("B,A,C,grids,img,conf_thres,conf_mode", [
    (3, 3, 20, [20, 10, 5], 160, -0.0151, "uniform"),        # N = 1 575  -> 1 CTA / image
    (2, 3, 80, [40, 20, 10], 320, -0.0151, "uniform"),       # N = 6 300  -> 1 CTA, nearly full
    (2, 3, 80, [52, 26, 13], 416, -0.0151, "uniform"),       # N = 10 647 -> cluster of 2
    (3, 3, 80, [80, 40, 20], 640, -0.0151, "uniform"),       # N = 25 200 -> cluster of 4 (headline shape)
    (2, 3, 5, [104, 52, 26], 832, -0.0151, "uniform"),       # N = 42 588 -> cluster of 8
    (2, 3, 300, [20, 10, 5], 160, -0.0151, "uniform"),       # C > 256   -> two class passes
    (2, 3, 80, [80, 40, 20], 640, 0.25, "sparse"),           # sparse survivors: most CTAs of the cluster idle
    (2, 5, 20, [13], 416, -0.0151, "uniform"),               # YOLOv2 odd grid
])
def test_cluster_sort_equals_lookback_sort(B, A, C, grids, img, conf_thres, conf_mode):
    levels = [t.to(DEV) for t in synth.yolo_planar(B=B, A=A, C=C, grids=grids, img=img, seed=77, conf_mode=conf_mode)]
    got = _run_yolo(levels, A, conf_thres, None)
    want = _run_yolo(levels, A, conf_thres, "global")
    assert sum(r.shape[0] for r, _ in want) > 0
    for b in range(B):
        assert got[b][0].shape == want[b][0].shape, f"image {b}: kept {got[b][0].shape[0]} vs {want[b][0].shape[0]}"
        assert torch.equal(got[b][1], want[b][1]), f"image {b}: kept candidate indices differ"
        assert torch.equal(got[b][0], want[b][0]), f"image {b}: rows differ"


@pytest.mark.parametrize("P,C,B", [(8732, 21, 4), (3000, 5, 3)])
def test_cluster_sort_prior_path(P, C, B):
    g = torch.Generator().manual_seed(5)
    pri = torch.rand(P, 4, generator=g) * 0.5 + 0.1
    loc = torch.randn(B, P, 4, generator=g) * 0.2
    cls = torch.randn(B, P, C, generator=g) * 2 - 1
    outs = []
    for mode in (None, "global"):
        with _sort_mode(mode):
            rows, index, count = prior_nms_raw(loc.to(DEV), cls.to(DEV), pri.to(DEV), topk=100, want_index=True)
            torch.cuda.synchronize()
        outs.append((rows.cpu(), index.cpu(), count.cpu()))
    k = outs[0][2][0].tolist()
    assert torch.equal(outs[0][2], outs[1][2])
    for b in range(B):
        assert torch.equal(outs[0][0][b, :k[b]], outs[1][0][b, :k[b]])
        assert torch.equal(outs[0][1][b, :k[b]], outs[1][1][b, :k[b]])


@pytest.mark.parametrize("case", ["sparse", "all_survive", "mixed", "just_over_capacity"])
def test_dense_route_for_large_images_equals_lookback_sort(case):
    """Images with more slots than one cluster sorts (> 53 248) take the dense route: compaction + cluster of 4, with the
    multi-launch sort gated per image for those whose survivors exceed the cluster (26 624).  All three situations —
    nobody overflows, everybody overflows, some do — must match the plain multi-launch sort bit for bit."""
    B, A, C, grids, img = 3, 3, 5, [160, 80, 40], 1280                      # N = 100 800, 101 376 slots
    if case == "sparse":
        lv, thr = synth.yolo_crowd(B, A, C, grids, img, seed=5), 0.001      # ~10 k survivors per image
    elif case == "all_survive":
        lv, thr = synth.yolo_crowd(B, A, C, grids, img, seed=6), -0.0151    # 100 800 survivors: every image overflows
    elif case == "mixed":
        lv, thr = synth.yolo_crowd(B, A, C, grids, img, seed=7), 0.001
        p = lv[0].view(B, A, 5 + C, 160, 160)
        p[1, :, 4] = p[1, :, 4].clamp_min(0.002)                            # image 1: all 76 800 cells of the first level survive
        synth.make_tie_free(lv, A)
    else:
        lv, thr = synth.yolo_crowd(B, A, C, grids, img, seed=8, keep_frac=0.27), 0.001   # ~27.2 k: around the capacity
    levels = [t.to(DEV) for t in lv]
    got = _run_yolo(levels, A, thr, None)
    want = _run_yolo(levels, A, thr, "global")
    ks = [r.shape[0] for r, _ in want]
    assert min(ks) > 0
    for b in range(B):
        assert got[b][0].shape == want[b][0].shape, f"{case} image {b}: kept {got[b][0].shape[0]} vs {want[b][0].shape[0]}"
        assert torch.equal(got[b][1], want[b][1]), f"{case} image {b}: kept candidate indices differ"
        assert torch.equal(got[b][0], want[b][0]), f"{case} image {b}: rows differ"

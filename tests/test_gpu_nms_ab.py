"""GPU: the two pair pre-filters of the class-aware NMS kernel — pair-wise half2 bound test (B200DET_NMS=half2) and
quantised bound tables + queued exact tests (B200DET_NMS=tab) — decide the SAME pairs with the same exact fp32 test, so
the whole pipeline must be bit-identical between them: kept rows, merged boxes, candidate indices, counts.  Covers every
kernel instantiation (128 / 256 / 512 threads), multi-chunk segments, and inputs that stress the quantisation: identical
boxes, far outliers (the level range collapses), non-finite coordinates, negative and sub-pixel coordinates."""
import os

import pytest
import torch

from objectdetectionpl_b200 import synth
from objectdetectionpl_b200.postprocess import yolo_nms_raw

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class _nms_mode:
    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        self.old = os.environ.get("B200DET_NMS")
        os.environ["B200DET_NMS"] = self.mode

    def __exit__(self, *a):
        if self.old is None:
            os.environ.pop("B200DET_NMS", None)
        else:
            os.environ["B200DET_NMS"] = self.old


def _run(levels, A, conf_thres, mode, nms_thres=0.4):
    with _nms_mode(mode):
        rows, index, count = yolo_nms_raw(levels, A, conf_thres, nms_thres, want_index=True)
        torch.cuda.synchronize()
    c = count.cpu().tolist()
    rows, index = rows.cpu(), index.cpu()
    return [(rows[b, :k].clone(), index[b, :k].clone()) for b, k in enumerate(c)]


def _same(levels, A, conf_thres=-0.0151, nms_thres=0.4):
    levels = [t.to(DEV) for t in levels]
    a = _run(levels, A, conf_thres, "half2", nms_thres)
    b = _run(levels, A, conf_thres, "tab", nms_thres)
    assert len(a) == len(b)
    for i, ((ra, ia), (rb, ib)) in enumerate(zip(a, b)):
        assert ra.shape == rb.shape, f"image {i}: {ra.shape[0]} rows (half2) vs {rb.shape[0]} (tab)"
        assert torch.equal(ia, ib), f"image {i}: kept candidates differ"
        assert torch.equal(ra.view(torch.int32), rb.view(torch.int32)), f"image {i}: rows differ bitwise"
    return a


@pytest.mark.parametrize("B,A,C,grids,img,thr,crowd", [
    (4, 3, 80, [80, 40, 20], 640, -0.0151, False),      # headline shape: 256 threads, ~315-row segments
    (3, 3, 80, [13, 26, 52], 416, -0.0151, False),      # YOLOv3: 128 threads, 192-row chunks
    (2, 3, 5, [160, 80, 40], 1280, 0.001, True),        # crowd: 512 threads, multi-chunk segments
    (2, 3, 1, [40, 20, 10], 320, -0.0151, False),       # one class: 6 300-row segment, 17 chunks
    (2, 5, 20, [13], 416, 0.3, False),                  # YOLOv2, filtered
    (3, 3, 4, [8, 4, 2], 64, -0.0151, False),           # tiny
])
def test_tab_prefilter_equals_half2_prefilter(B, A, C, grids, img, thr, crowd):
    lv = synth.yolo_crowd(B, A, C, grids, img, seed=5) if crowd else synth.yolo_planar(B, A, C, grids, img, 41, v5_view=False)
    out = _same(lv, A, thr)
    assert sum(r.shape[0] for r, _ in out) > 0


@pytest.mark.parametrize("nms_thres", [0.0, 0.05, 0.4, 0.9, 0.999])
def test_tab_prefilter_threshold_sweep(nms_thres):
    lv = synth.yolo_planar(2, 3, 3, [20, 10, 5], 160, 43)
    _same(lv, 3, nms_thres=nms_thres)


def _planar(B, A, C, G):
    return torch.zeros(B, A, 5 + C, G, G)


def test_tab_prefilter_adversarial_geometry():
    g = torch.Generator().manual_seed(9)
    B, A, C, G = 2, 3, 2, 16
    N = A * G * G
    base = synth.yolo_planar(B, A, C, [G], 128, 44)[0].view(B, A, 5 + C, G, G)
    cases = {}
    t = base.clone(); t[:, :, 0:2] = 50.0; t[:, :, 2:4] = 20.0
    cases["identical boxes"] = t
    t = base.clone(); t[0, 0, 0, 0, 0] = 3.0e9; t[0, 1, 1, 3, 3] = -2.5e12; t[1, 2, 2, 5, 5] = 1.0e30
    cases["far outliers collapse the level range"] = t
    t = base.clone(); t[0, 0, 0, 1, 1] = float("nan"); t[0, 1, 2, 2, 2] = float("inf"); t[1, 0, 1, 0, 0] = float("-inf")
    t[1, 1, 3, 4, 4] = float("nan")
    cases["non-finite coordinates"] = t
    t = base.clone(); t[:, :, 0:2] = torch.rand(B, A, 2, G, G, generator=g) * 4 - 2; t[:, :, 2:4] = torch.rand(B, A, 2, G, G, generator=g) * 0.5
    cases["sub-pixel boxes around the origin"] = t
    t = base.clone(); t[:, :, 2:4] = -torch.rand(B, A, 2, G, G, generator=g) * 0.9
    cases["negative widths above -1"] = t
    t = base.clone(); t[:, :, 0] = 64.0; t[:, :, 1] = torch.arange(G * G, dtype=torch.float32).view(G, G) * 0.37; t[:, :, 2] = 30.0; t[:, :, 3] = 9.0
    cases["a dense vertical chain"] = t
    for name, t in cases.items():
        synth.make_tie_free([t], A)
        try:
            _same([t.reshape(B, A * (5 + C), G, G)], A)
        except AssertionError as e:
            raise AssertionError(f"{name}: {e}")

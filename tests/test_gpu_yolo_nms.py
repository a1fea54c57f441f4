"""GPU parity: the CUDA YOLO post-processing pipeline (through the C ABI) against the golden vectors of
the unmodified reference and against the CPU oracle on seeded inputs.
Bar: kept candidate indices, obj conf, class conf and class id bit-exact; merged boxes within 1e-5 rel."""
import ctypes

import numpy as np
import pytest
import torch

import objectdetectionpl_b200 as od
from objectdetectionpl_b200 import _lib as L, synth
from oracle import ref_port as rp
from tests.golden_io import load, unpack_list, yolo_levels, YOLO_CASES, assert_rows_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cuda(levels):
    return [t.to(DEV) for t in levels]


def _check(got, gidx, want, widx, what):
    assert len(got) == len(want)
    for b in range(len(want)):
        if want[b] is None:
            assert got[b] is None, f"{what}[{b}]"
            continue
        assert got[b] is not None, f"{what}[{b}]: nothing kept"
        if widx is not None:
            gi, wi = gidx[b].cpu(), widx[b]
            assert gi.shape == wi.shape, f"{what}[{b}]: kept {gi.shape[0]} vs {wi.shape[0]}"
            bad = (gi != wi).nonzero().flatten()
            assert bad.numel() == 0, f"{what}[{b}]: kept index differs first at row {bad[:5].tolist()}"
        assert_rows_close(got[b], want[b], rtol=1e-5, atol=1e-4, what=f"{what}[{b}]")


@pytest.mark.parametrize("name", YOLO_CASES)
def test_golden_reference_vectors(name):
    d = load(name)
    levels = yolo_levels(d)
    want = unpack_list(d, "out")
    A = int(d["A"])
    fn = od.non_max_suppression_v2 if A == 5 else od.non_max_suppression
    arg = _cuda(levels)
    if len(arg) == 1:
        arg = arg[0]                     # single-tensor branch (model/YOLOV3.py:281-282)
    got = fn(None, arg)
    _check(got, None, want, None, name)
    # the kept candidate indices against the oracle
    got, gidx = fn(None, _cuda(levels), return_index=True)
    want2, widx = rp.yolo_nms(levels, num_anchors=A, return_index=True)
    _check(got, gidx, want2, widx, name + "/oracle")


@pytest.mark.parametrize("B,A,C,grids,img,seed", [
    (3, 3, 20, [20, 10, 5], 160, 101),          # N=1575
    (2, 3, 80, [40, 20, 10], 320, 102),         # N=6300, COCO classes
    (1, 3, 2, [24, 12, 6], 192, 103),           # 2 classes -> ~1100-row segments (multi-chunk NMS)
    (2, 5, 20, [13], 416, 104),                 # YOLOv2, odd grid -> scalar load path
    (2, 3, 1, [16, 8], 128, 105),               # single class
    (1, 3, 300, [8, 4], 64, 106),               # >256 classes -> two class passes
])
def test_against_oracle(B, A, C, grids, img, seed):
    levels = synth.yolo_planar(B, A, C, grids, img, seed)
    fn = od.non_max_suppression_v2 if A == 5 else od.non_max_suppression
    got, gidx = fn(None, _cuda(levels), return_index=True)
    want, widx = rp.yolo_nms_fast(rp.yolo_rows_from_planar(levels, A))
    _check(got, gidx, want, widx, f"seed{seed}")


def test_dense_crowd_multichunk_conf_threshold():
    """config-5-like: clustered boxes, compat=False with conf_thres=0.001, few classes -> deep suppression
    chains and segments far longer than one 512-row chunk."""
    levels = synth.yolo_crowd(B=2, A=3, C=2, grids=[40, 20], img=320, seed=5, blobs=12, keep_frac=0.5)
    got, gidx = od.non_max_suppression(None, _cuda(levels), conf_thres=0.001, compat=False, return_index=True)
    want, widx = rp.yolo_nms_fast(rp.yolo_rows_from_planar(levels, 3), conf_thres=0.001)
    _check(got, gidx, want, widx, "crowd")


def test_threshold_filters_everything_and_nothing():
    levels = synth.yolo_planar(2, 3, 4, [8, 4], 64, 9)
    got = od.non_max_suppression(None, _cuda(levels), conf_thres=2.0, compat=False)
    assert got == [None, None]
    got_c = od.non_max_suppression(None, _cuda(levels), conf_thres=2.0)       # compat: threshold ignored
    assert all(g is not None for g in got_c)


def test_nms_thres_and_v5_view_equivalence():
    lv5 = synth.yolo_planar(2, 3, 6, [16, 8, 4], 128, 33, v5_view=True)
    lv3 = [t.reshape(t.shape[0], -1, t.shape[2], t.shape[3]) for t in lv5]
    a = od.non_max_suppression(None, _cuda(lv5), nms_thres=0.6)
    b = od.non_max_suppression(None, _cuda(lv3), nms_thres=0.6)
    want = rp.yolo_nms_rows(rp.yolo_rows_from_planar(lv5, 3), nms_thres=0.6)
    for i in range(2):
        assert torch.equal(a[i], b[i])
        assert_rows_close(a[i], want[i], rtol=1e-5, atol=1e-4, what=f"nms0.6[{i}]")


def test_score_ties_ordered_by_candidate_index():
    """Published tie rule: equal scores keep ascending candidate order (== argsort(stable=True))."""
    levels = synth.yolo_planar(1, 3, 3, [8], 64, 77, tie_free=False)
    p = levels[0].view(1, 3, 8, 8, 8)
    p[:, :, 4] = 0.5                      # same conf everywhere
    p[:, :, 5:] = 0.0
    p[:, :, 5] = 0.25                     # same class, same class conf -> all scores tie
    got, gidx = od.non_max_suppression(None, _cuda(levels), return_index=True)
    want, widx = rp.yolo_nms_rows(rp.yolo_rows_from_planar(levels, 3), return_index=True)
    assert torch.equal(gidx[0].cpu(), widx[0])
    assert_rows_close(got[0], want[0], rtol=1e-5, atol=1e-4, what="ties")


def test_stage_outputs_sorted_and_segmented():
    """White-box: after the sort stage every image is ordered by (class asc, score desc, index asc) and
    seg_off delimits the classes."""
    lib = L.load()
    B, A, C, grids = 2, 3, 7, [16, 8]
    levels = _cuda(synth.yolo_planar(B, A, C, grids, 128, 55))
    from objectdetectionpl_b200.postprocess import _yolo_desc
    d = _yolo_desc(levels, A, -0.0151, 0.4, None, None, None)
    nbytes = lib.b200det_yolo_workspace_bytes(ctypes.byref(d))
    ws = torch.zeros(nbytes, dtype=torch.uint8, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    L.check(lib.b200det_yolo_stage_reset(ctypes.byref(d), ws.data_ptr(), nbytes, st))
    L.check(lib.b200det_yolo_stage_decode(ctypes.byref(d), ws.data_ptr(), nbytes, st))
    L.check(lib.b200det_yolo_stage_sort(ctypes.byref(d), ws.data_ptr(), nbytes, st))
    torch.cuda.synchronize()

    def field(name, dtype):
        off, nb = ctypes.c_size_t(), ctypes.c_size_t()
        L.check(lib.b200det_yolo_workspace_field(ctypes.byref(d), name.encode(), ctypes.byref(off), ctypes.byref(nb)))
        return ws[off.value:off.value + nb.value].view(dtype).cpu()

    N = sum(A * g * g for g in grids)
    n_pad = sum((A * g * g + 511) // 512 * 512 for g in grids)      # every level starts on a tile boundary
    count = field("count", torch.int32)
    assert count.tolist() == [N] * B
    pay = field("sorted_pay", torch.int32).view(B, n_pad)
    rank = field("sorted_rank", torch.int32).view(B, n_pad)
    seg = field("seg_off", torch.int32).view(B, C + 1)
    orig = field("orig", torch.int32).view(B, n_pad)               # slot -> candidate index (levels start on tile boundaries)
    rows = rp.yolo_rows_from_planar([t.cpu() for t in levels], A)
    for b in range(B):
        cls_conf, cls_id = rows[b, :, 5:].max(1)
        score = rows[b, :, 4] * cls_conf
        order = torch.argsort(-score, stable=True)                 # global score rank -> candidate
        key = cls_id[order] * (N + 1) + torch.arange(N)            # stable partition by class
        want_pos = torch.argsort(key, stable=True)                 # sorted position -> rank
        slot = pay[b, :N] & 0xFFFFF
        assert torch.equal(rank[b, :N].long(), want_pos), "rank of sorted positions"
        assert torch.equal(orig[b][slot.long()].long(), order[want_pos]), "candidate at sorted positions"
        assert torch.equal((pay[b, :N] >> 20).long(), cls_id[order[want_pos]])
        hist = torch.bincount(cls_id, minlength=C)
        assert torch.equal(seg[b].long(), torch.cat([torch.zeros(1, dtype=torch.long), hist.cumsum(0)]))


def test_no_cpu_fallback():
    levels = synth.yolo_planar(1, 3, 2, [4], 32, 1)
    with pytest.raises(RuntimeError, match="no CPU path"):
        od.non_max_suppression(None, levels)


def test_tma_variant_of_k1_is_bit_identical(monkeypatch):
    """The bulk-async (cp.async.bulk + mbarrier) pipeline variant of the decode kernel (B200DET_K1=tma) must produce
    exactly the candidates of the default LDG kernel, in all decode modes."""
    levels = _cuda(synth.yolo_planar(3, 3, 20, [40, 20, 12], 320, 41))          # G*G = 1600, 400, 144: 1-4 runs per tile
    a, ai = od.non_max_suppression(None, levels, return_index=True)
    monkeypatch.setenv("B200DET_K1", "tma")
    b, bi = od.non_max_suppression(None, levels, return_index=True)
    for x, y, xi, yi in zip(a, b, ai, bi):
        assert torch.equal(x, y) and torch.equal(xi, yi)
    heads = [synth.raw_logits(2, 3, 6, G, 80 + G).to(DEV) for G in (16, 12)]
    for h in heads:
        h.view(2, 3, 11, h.shape[2], h.shape[3])[:, :, 4] += 3.0
    kw = dict(conf_thres=0.25, compat=False, decode="yolov5", strides=[8.0, 16.0],
              anchors=[torch.tensor([[10., 13.], [16., 30.], [33., 23.]]), torch.tensor([[30., 61.], [62., 45.], [59., 119.]])])
    t = od.non_max_suppression(None, heads, **kw)
    monkeypatch.delenv("B200DET_K1")
    l = od.non_max_suppression(None, heads, **kw)
    for x, y in zip(t, l):
        assert torch.equal(x, y)


def test_pipeline_is_cuda_graph_capturable():
    """The library never allocates or synchronises (include/b200det.h), so the four launches of a step can be captured
    once and replayed; the replay must reproduce the eager result for new head values in the same buffers."""
    levels = _cuda(synth.yolo_planar(4, 3, 20, [20, 10, 5], 160, 55))
    static = [t.clone() for t in levels]
    od.yolo_nms_raw(static, 3, want_index=True)                       # warm-up: one-time function attributes, workspace
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        rows, index, count = od.yolo_nms_raw(static, 3, want_index=True)
    for seed in (56, 57):
        fresh = _cuda(synth.yolo_planar(4, 3, 20, [20, 10, 5], 160, seed))
        for dst, src in zip(static, fresh):
            dst.copy_(src)
        g.replay()
        torch.cuda.synchronize()
        want_rows, want_index, want_count = od.yolo_nms_raw(fresh, 3, want_index=True)
        torch.cuda.synchronize()
        assert torch.equal(count, want_count)
        for b, k in enumerate(count.tolist()):
            assert torch.equal(rows[b, :k], want_rows[b, :k]) and torch.equal(index[b, :k], want_index[b, :k])

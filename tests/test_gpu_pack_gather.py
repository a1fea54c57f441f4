"""GPU: the packed result form of the pipeline, the device-side detection packing and the exchange step's host logic
(world size 1 here; the 2-rank gloo logic is in tests/test_dist_gloo.py, the N-GPU NCCL run in tests/test_gpu_multi.py)."""
import ctypes

import pytest
import torch

import objectdetectionpl_b200 as od
from objectdetectionpl_b200 import _lib as L, dist as odist, synth
from objectdetectionpl_b200.postprocess import _yolo_desc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("B,A,C,grids,img,thr", [(5, 3, 6, [20, 10, 5], 160, None), (3, 3, 4, [13, 26], 208, None),
                                                 (4, 3, 5, [16, 8], 128, 0.6), (2, 3, 3, [8], 64, 2.0)])
def test_packed_pipeline_equals_padded_pipeline(B, A, C, grids, img, thr):
    lv = [t.to(DEV) for t in synth.yolo_planar(B, A, C, grids, img, 31, v5_view=True)]
    kw = dict(compat=False, conf_thres=thr) if thr is not None else {}
    rows, index, count = od.yolo_nms_raw(lv, A, od.YOLO_FORCED_CONF_THRES if thr is None else thr, want_index=True)
    got, gidx = od.non_max_suppression(None, lv, return_index=True, **kw)                 # packed form underneath
    counts = count.cpu().tolist()
    assert len(got) == B
    for b, k in enumerate(counts):
        if k == 0:
            assert got[b] is None and gidx[b] is None
            continue
        assert torch.equal(got[b], rows[b, :k]) and torch.equal(gidx[b], index[b, :k].long())
        assert got[b].is_contiguous()
    if thr == 2.0:
        assert all(g is None for g in got)                                                # nothing survives conf >= 2
    # a second call with the same shapes re-uses the plan (other data, other pointers)
    lv2 = [t.to(DEV) for t in synth.yolo_planar(B, A, C, grids, img, 32, v5_view=True)]
    again = od.non_max_suppression(None, lv2, **kw)
    rows2, _, count2 = od.yolo_nms_raw(lv2, A, od.YOLO_FORCED_CONF_THRES if thr is None else thr)
    for b, k in enumerate(count2.cpu().tolist()):
        assert (again[b] is None) if k == 0 else torch.equal(again[b], rows2[b, :k])


def test_packed_c_abi_offsets_and_canaries():
    """b200det_yolo_nms_packed through ctypes: offsets = exclusive prefix of the counts, rows/index packed back to back,
    nothing written past offsets[B] (canary), same rows as the padded entry."""
    lib = L.load()
    B, A, C = 6, 3, 5
    lv = [t.to(DEV) for t in synth.yolo_planar(B, A, C, [16, 8, 4], 128, 77, v5_view=True)]
    d = _yolo_desc(lv, A, 0.3, 0.4, None, None, None)
    n, n_pad = ctypes.c_int32(), ctypes.c_int32()
    L.check(lib.b200det_yolo_num_candidates(ctypes.byref(d), ctypes.byref(n), ctypes.byref(n_pad)))
    wsb = lib.b200det_yolo_workspace_bytes(ctypes.byref(d))
    ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
    cap = B * n_pad.value
    rows = torch.full((cap, 7), -7.0, device=DEV)
    idx = torch.full((cap,), -7, dtype=torch.int32, device=DEV)
    cnt = torch.empty(B, dtype=torch.int32, device=DEV)
    off = torch.empty(B + 1, dtype=torch.int32, device=DEV)
    st = L.stream_ptr(torch.device(DEV))
    early = torch.full((2 * B + 1,), -1, dtype=torch.int32).pin_memory()       # written by the device (mapped pinned memory)
    ev = torch.cuda.Event()
    ev.record()
    L.check(lib.b200det_yolo_nms_packed(ctypes.byref(d), ws.data_ptr(), wsb, rows.data_ptr(), idx.data_ptr(), cnt.data_ptr(),
                                        off.data_ptr(), early.data_ptr(), ev.cuda_event, st))
    ev.synchronize()                                                            # counts are final before the rows are
    early_now = early.clone()
    prow, pidx, pcnt = od.yolo_nms_raw(lv, A, 0.3, want_index=True)
    torch.cuda.synchronize()
    c, o = cnt.cpu(), off.cpu()
    assert torch.equal(early_now[:B], c) and torch.equal(early_now[B:], o)
    assert torch.equal(c, pcnt.cpu()) and o[0] == 0 and torch.equal(o[1:], torch.cumsum(c, 0).int())
    total = int(o[B])
    assert 0 < total < cap
    for b in range(B):
        assert torch.equal(rows[o[b]:o[b + 1]], prow[b, :c[b]]) and torch.equal(idx[o[b]:o[b + 1]], pidx[b, :c[b]])
    assert bool((rows[total:] == -7.0).all()) and bool((idx[total:] == -7).all())


def test_pack_detections_raw_matches_list_packing():
    B, A, C = 7, 3, 4
    lv = [t.to(DEV) for t in synth.yolo_planar(B, A, C, [10, 5], 80, 5, v5_view=True)]
    rows, _, count = od.yolo_nms_raw(lv, A, 0.5)                              # some images lose rows at conf >= 0.5
    dets = od.non_max_suppression(None, lv, conf_thres=0.5, compat=False)
    want, wc = odist._pack_list(dets, image_offset=40, device=torch.device(DEV))
    out = torch.full((B * rows.shape[1] + 3, 8), 9.5, device=DEV)
    packed, offsets = odist.pack_detections_raw(rows, count, 40, out=out)
    off = offsets.cpu()
    assert torch.equal(off[1:] - off[:-1], count.cpu()) and int(off[B]) == want.shape[0]
    assert torch.equal(packed[:want.shape[0]], want)
    assert bool((out[want.shape[0]:] == 9.5).all())                           # nothing written past the last row
    # capacity smaller than the total: rows beyond it are dropped, the total is still reported
    small = torch.full((want.shape[0] // 2, 8), 9.5, device=DEV)
    p2, off2 = odist.pack_detections_raw(rows, count, 40, out=small)
    assert int(off2[B]) == want.shape[0] and torch.equal(p2, want[: small.shape[0]])
    # the exchange step with one rank: slicing by offsets gives the per-image list back
    g = odist.gather_detections_raw(rows, count, 40)
    per = g.per_image()
    assert len(per) == B and g.totals == [want.shape[0]]
    for a, b in zip(per, dets):
        assert (a is None and b is None) or torch.equal(a, b)
    assert torch.equal(g.packed(), want)
    back = odist.unpack_detections(want, 40 + B)
    for i in range(B):
        assert (back[40 + i] is None and dets[i] is None) or torch.equal(back[40 + i], dets[i])

"""GPU: out-of-bounds guards.  compute-sanitizer is not available on this pool, so every buffer the C ABI writes is
wrapped in canary zones: the workspace (the library lays ~20 arrays out in it), out_rows / out_index / out_count (padded
and packed forms) and the prior pipeline's buffers.  A kernel that writes one element outside its buffer trips a canary."""
import ctypes

import pytest
import torch

from objectdetectionpl_b200 import _lib as L, synth
from objectdetectionpl_b200.postprocess import _yolo_desc

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
GUARD = 1 << 16          # bytes on each side
PAT = 0x5A


class Guarded:
    """`nbytes` usable bytes with GUARD canary bytes before and after (256-byte aligned start)."""

    def __init__(self, nbytes):
        self.n = int(nbytes)
        self.buf = torch.full((self.n + 2 * GUARD + 256,), PAT, dtype=torch.uint8, device=DEV)
        base = self.buf.data_ptr()
        self.off = (-(base + GUARD)) % 256 + GUARD
        self.ptr = base + self.off

    def view(self, dtype, shape):
        return self.buf[self.off:self.off + self.n].view(dtype).view(shape)

    def check(self, what):
        lo, hi = self.buf[:self.off], self.buf[self.off + self.n:]
        assert bool((lo == PAT).all()), f"{what}: bytes BEFORE the buffer were written"
        assert bool((hi == PAT).all()), f"{what}: bytes AFTER the buffer were written"


YOLO_CASES = [
    # B, A, C, grids, img, conf_thres, crowd
    (3, 3, 80, [80, 40, 20], 640, -0.0151, False),       # headline shapes: cluster sort, 384-row chunks
    (5, 3, 7, [13, 26, 52], 416, -0.0151, False),        # odd plane (scalar K1 tiles), short segments
    (2, 5, 20, [13], 416, 0.3, False),                   # YOLOv2: 5 anchors, one level, some rows filtered
    (2, 3, 5, [160, 80, 40], 1280, 0.001, True),         # crowd: look-back sort, multi-chunk segments, 512-thread NMS
    (2, 3, 300, [8, 4], 64, -0.0151, False),             # > 256 classes: two class passes
]


@pytest.mark.parametrize("B,A,C,grids,img,thr,crowd", YOLO_CASES)
@pytest.mark.parametrize("packed", [False, True])
def test_yolo_pipeline_writes_stay_inside_their_buffers(B, A, C, grids, img, thr, crowd, packed):
    lib = L.load()
    lv = synth.yolo_crowd(B, A, C, grids, img, seed=5) if crowd else synth.yolo_planar(B, A, C, grids, img, 17, tie_free=False)
    lv = [t.to(DEV) for t in lv]
    d = _yolo_desc(lv, A, thr, 0.4, None, None, None)
    n, n_pad = ctypes.c_int32(), ctypes.c_int32()
    L.check(lib.b200det_yolo_num_candidates(ctypes.byref(d), ctypes.byref(n), ctypes.byref(n_pad)))
    P = n_pad.value
    wsb = lib.b200det_yolo_workspace_bytes(ctypes.byref(d))
    ws, rows, idx = Guarded(wsb), Guarded(B * P * 7 * 4), Guarded(B * P * 4)
    cnt, off = Guarded(B * 4), Guarded((B + 1) * 4)
    st = L.stream_ptr(DEV)
    for _ in range(2):                                    # twice: the second run starts from a used workspace
        if packed:
            L.check(lib.b200det_yolo_nms_packed(ctypes.byref(d), ws.ptr, wsb, rows.ptr, idx.ptr, cnt.ptr, off.ptr, None, None, st))
        else:
            L.check(lib.b200det_yolo_nms(ctypes.byref(d), ws.ptr, wsb, rows.ptr, idx.ptr, cnt.ptr, st))
    torch.cuda.synchronize()
    for g, nm in ((ws, "workspace"), (rows, "out_rows"), (idx, "out_index"), (cnt, "out_count"), (off, "out_offsets")):
        g.check(f"{nm} ({'packed' if packed else 'padded'})")
    c = cnt.view(torch.int32, (B,)).cpu()
    assert int(c.min()) >= 0 and int(c.max()) <= n.value
    if packed:
        o = off.view(torch.int32, (B + 1,)).cpu()
        assert torch.equal(o[1:] - o[:-1], c) and int(o[0]) == 0
        tail = rows.view(torch.uint8, (B * P * 7 * 4,))[int(o[B]) * 28:]
        assert bool((tail == PAT).all()), "packed rows written past the total"
    else:
        r = rows.view(torch.uint8, (B, P * 28))
        for b in range(B):
            assert bool((r[b, int(c[b]) * 28:] == PAT).all()), f"image {b}: rows written past its count"


@pytest.mark.parametrize("which,B,C,compat", [("ssd", 4, 80, 1), ("ssd", 3, 5, 0), ("retina", 2, 80, 1), ("retina128", 3, 6, 1)])
def test_prior_pipeline_writes_stay_inside_their_buffers(which, B, C, compat):
    lib = L.load()
    pri = {"ssd": synth.ssd_priors, "retina": lambda: synth.retina_priors(800), "retina128": lambda: synth.retina_priors(128)}[which]()
    loc, cls = synth.prior_heads(B, pri.shape[0], C, 3, cls_mean=-3.0)
    pri, loc, cls = pri.to(DEV), loc.to(DEV), cls.to(DEV)
    d = L.PriorDesc()
    d.batch, d.num_priors, d.num_classes = B, pri.shape[0], C
    d.loc, d.cls, d.priors = loc.data_ptr(), cls.data_ptr(), pri.data_ptr()
    d.topk, d.nms_thresh, d.class_thresh, d.mode_min, d.compat = 100, 0.5, 0.45, 0, compat
    wsb = lib.b200det_prior_workspace_bytes(ctypes.byref(d))
    ws, rows, idx = Guarded(wsb), Guarded(B * 100 * 7 * 4), Guarded(B * 100 * 4)
    cnt, cand = Guarded(B * 4), Guarded(B * 4)
    for _ in range(2):
        L.check(lib.b200det_prior_nms(ctypes.byref(d), ws.ptr, wsb, rows.ptr, idx.ptr, cnt.ptr, cand.ptr, L.stream_ptr(DEV)))
    torch.cuda.synchronize()
    for g, nm in ((ws, "workspace"), (rows, "out_rows"), (idx, "out_index"), (cnt, "out_count"), (cand, "cand_count")):
        g.check(nm)
    c = cnt.view(torch.int32, (B,)).cpu()
    assert int(c.min()) >= 0 and int(c.max()) <= 100
    r = rows.view(torch.uint8, (B, 100 * 28))
    for b in range(B):
        assert bool((r[b, int(c[b]) * 28:] == PAT).all()), f"image {b}: rows written past its count"


def test_pack_and_d3_writes_stay_inside_their_buffers():
    lib = L.load()
    B, A, C, G = 3, 3, 6, 20
    lv = [t.to(DEV) for t in synth.yolo_planar(B, A, C, [G, G // 2], 160, 23)]
    import objectdetectionpl_b200 as od
    rows, _, count = od.yolo_nms_raw(lv, A, 0.4)
    total = int(count.sum())
    out, offs = Guarded(total * 32), Guarded((B + 1) * 4)
    L.check(lib.b200det_pack_detections(rows.data_ptr(), count.data_ptr(), B, rows.shape[1], rows.shape[1], 0, out.ptr, total,
                                        offs.ptr, L.stream_ptr(DEV)))
    head = synth.raw_logits(B, A, C, G, 4).to(DEV)
    N = A * G * G
    anc = torch.tensor([[1.5, 2.0], [3.0, 2.5], [4.0, 6.0]], device=DEV)
    boxes, confs, det = Guarded(B * N * 16), Guarded(B * N * C * 4), Guarded(B * N * 4)
    L.check(lib.b200det_yolo_forward_dynamic(head.data_ptr(), B, A, C, G, G, anc.data_ptr(), 1.05, boxes.ptr, 4, confs.ptr, C,
                                             det.ptr, 1, L.stream_ptr(DEV)))
    torch.cuda.synchronize()
    for g, nm in ((out, "packed rows"), (offs, "offsets"), (boxes, "d3 boxes"), (confs, "d3 confs"), (det, "d3 det")):
        g.check(nm)

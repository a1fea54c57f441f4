"""GPU: layout='channels_last' (heads stored [B, A, G, G, 5+C]) gives exactly the detections of the
planar path on the permuted tensor — the planar path being the one pinned against the reference."""
import pytest
import torch

import objectdetectionpl_b200 as od
from objectdetectionpl_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _to_channels_last(levels, A):
    out = []
    for t in levels:
        B, G = t.shape[0], t.shape[2]
        F = t.numel() // (B * A * G * G)
        out.append(t.reshape(B, A, F, G, G).permute(0, 1, 3, 4, 2).contiguous())
    return out


@pytest.mark.parametrize("B,A,C,grids,img,seed", [
    (2, 3, 80, [40, 20, 10], 320, 1),        # 5+C = 85 odd: conflict-free row walk, aligned blocks
    (3, 3, 20, [20, 10, 5], 160, 2),         # level boundaries inside tiles and inside 32-row chunks
    (2, 3, 5, [16, 8], 128, 3),              # 5+C = 10 even
    (2, 5, 20, [13], 416, 4),                # odd grid: misaligned blocks take the scalar path
    (1, 3, 1, [12, 6], 96, 5),               # one class
])
def test_channels_last_equals_planar(B, A, C, grids, img, seed):
    planar = synth.yolo_planar(B=B, A=A, C=C, grids=grids, img=img, seed=seed)
    fn = od.non_max_suppression_v2 if A == 5 else od.non_max_suppression
    want, widx = fn(None, [t.to(DEV) for t in planar], return_index=True)
    cl = [t.to(DEV) for t in _to_channels_last(planar, A)]
    if A == 5:
        rows, index, count = od.yolo_nms_raw(cl, 5, want_index=True, layout="channels_last")
        got = [rows[b, :k] for b, k in enumerate(count.tolist())]
        gidx = [index[b, :k].long() for b, k in enumerate(count.tolist())]
    else:
        got, gidx = fn(None, cl, return_index=True, layout="channels_last")
    for b in range(B):
        assert torch.equal(gidx[b], widx[b]), f"image {b}: kept candidates differ"
        assert torch.equal(got[b], want[b]), f"image {b}: rows differ"


def test_channels_last_with_decode_and_threshold():
    heads = [synth.raw_logits(2, 3, 6, G, 50 + G) for G in (16, 8)]
    for h in heads:
        h.view(2, 3, 11, h.shape[2], h.shape[3])[:, :, 4] += 3.0
    kw = dict(conf_thres=0.25, compat=False, decode="yolov5", strides=[8.0, 16.0],
              anchors=[torch.tensor([[10., 13.], [16., 30.], [33., 23.]]), torch.tensor([[30., 61.], [62., 45.], [59., 119.]])])
    want = od.non_max_suppression(None, [h.to(DEV) for h in heads], **kw)
    got = od.non_max_suppression(None, [t.to(DEV) for t in _to_channels_last(heads, 3)], layout="channels_last", **kw)
    for g, w in zip(got, want):
        assert (g is None) == (w is None)
        if w is not None:
            assert torch.equal(g, w)


def test_channels_last_shape_is_checked():
    planar = synth.yolo_planar(B=1, A=3, C=4, grids=[8], img=64, seed=1)
    with pytest.raises(ValueError):
        od.non_max_suppression(None, [t.to(DEV) for t in planar], layout="channels_last")
    with pytest.raises(ValueError):
        od.non_max_suppression(None, [t.to(DEV) for t in planar], layout="nhwc")

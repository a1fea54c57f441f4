"""GPU parity: elementwise box maths (accuracy.py:6-114, 289-295) and decode_box (D1/D2)."""
import pytest
import torch

import objectdetectionpl_b200 as od
from objectdetectionpl_b200 import synth
from oracle import ref_port as rp
from tests.golden_io import load, T

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_iou_family_golden():
    d = load("iou")
    b1, b2, c1, c2 = (T(d[k]).to(DEV) for k in ("b1", "b2", "c1", "c2"))
    assert torch.equal(od.xywh2xyxy(b1).cpu(), T(d["c1"]))
    assert torch.equal(od.bbox_iou(b1, b2, x1y1x2y2=False).cpu(), T(d["plus1_xywh"]))
    assert torch.equal(od.bbox_iou(c1, c2).cpu(), T(d["plus1_xyxy"]))
    assert torch.equal(od.bbox_iou(c1[:1], c2).cpu(), T(d["plus1_one_vs_all"]))
    assert torch.equal(od.iou(c1.clamp(0, 100), c2.clamp(0, 100)).cpu(), T(d["pair_iou"]))
    n = b1.shape[0]
    gw = torch.linspace(0.5, 1.5, n, device=DEV)
    for kind in ("IoU", "GIoU", "DIoU", "CIoU"):
        for corner in (False, True):
            a = (c1 if corner else b1).t().clone().requires_grad_(True)     # transposed views, as losses.py:118 passes
            bb = (c2 if corner else b2).t()
            kw = {} if kind == "IoU" else {kind: True}
            v = od.bbox_iou_v5(a, bb, x1y1x2y2=corner, **kw)
            (v * gw).sum().backward()
            tag = f"v5_{kind}_{'xyxy' if corner else 'xywh'}"
            want, wgrad = T(d[tag]), T(d[tag + "_grad"])
            if kind == "CIoU":      # atan differs in the last ulp between libm/SLEEF and CUDA
                torch.testing.assert_close(v.detach().cpu(), want, rtol=1e-5, atol=1e-6, msg=tag)
            else:
                assert torch.equal(v.detach().cpu(), want), tag
            torch.testing.assert_close(a.grad.cpu(), wgrad, rtol=2e-4, atol=1e-6, msg=tag + " grad")


def test_bbox_iou_v5_contiguous_and_strided_agree():
    g = torch.Generator().manual_seed(3)
    p = torch.rand(500, 4, generator=g).to(DEV) + 0.1
    t = torch.rand(500, 4, generator=g).to(DEV) + 0.1
    a = od.bbox_iou_v5(p.t(), t.t(), x1y1x2y2=False, GIoU=True)
    b = od.bbox_iou_v5(p.t().contiguous(), t.t().contiguous(), x1y1x2y2=False, GIoU=True)
    assert torch.equal(a, b)
    want = rp.bbox_iou_v5(p.cpu().t(), t.cpu().t(), x1y1x2y2=False, GIoU=True)
    assert torch.equal(a.cpu(), want)


@pytest.mark.parametrize("mode", ["yolo_exp", "yolov5", "none"])
# 13, 7: odd planes -> 32 x 64 tile kernel; 20, 52, 6: 64-cell tile kernel (exact, ragged last tile, single short tile);
# C = 123 is the widest row the 64-cell tile holds, C = 124 falls back
@pytest.mark.parametrize("G,C", [(13, 4), (20, 80), (7, 130), (52, 80), (6, 123), (10, 124), (80, 1)])
def test_decode_box(mode, G, C):
    B, A = 2, 3
    head = synth.raw_logits(B, A, C, G, 61 + G)
    stride = 32.0
    if mode == "yolo_exp":
        anc = torch.tensor([[3.625, 2.8125], [4.875, 6.1875], [11.65625, 10.1875]])
        want = rp.decode_yolo_exp(head, anc, stride)
    elif mode == "yolov5":
        anc = torch.tensor([[116., 90.], [156., 198.], [373., 326.]])
        want = rp.decode_yolov5(head, anc, stride)
    else:
        anc = None
        want = rp.yolo_rows_from_planar([head], A)
    got = od.decode_box(head.to(DEV), anc, stride, mode, num_anchors=A).cpu()
    assert got.shape == want.shape
    if mode == "none":
        assert torch.equal(got, want)
    else:
        # 1e-5 relative; the absolute floor covers (2*sigmoid - 0.5 + g) * stride cancelling to ~0 in the first grid column,
        # where one ulp of the sigmoid (6e-8) times the stride (32) is all that is left of the value
        torch.testing.assert_close(got, want, rtol=1e-5, atol=4e-6 if mode == "yolov5" else 1e-6)


def test_decode_box_golden_d1():
    d = load("decode")
    head = T(d["head"])
    stride = int(d["img"]) / head.shape[2]
    got = od.decode_box(head.to(DEV), T(d["d1_scaled_anchors"]), stride, "yolo_exp").cpu()
    torch.testing.assert_close(got, T(d["d1_output"]), rtol=1e-5, atol=1e-6)


def test_fused_decode_nms_matches_decode_then_nms():
    """Extension path: non_max_suppression(decode='yolov5') == decode_box per level, then the plain NMS on the
    decoded rows (oracle NMS on the GPU-decoded values, so only the NMS is under test here)."""
    B, A, C = 2, 3, 6
    grids, strides = [16, 8], [8.0, 16.0]
    anchors = [torch.tensor([[10., 13.], [16., 30.], [33., 23.]]), torch.tensor([[30., 61.], [62., 45.], [59., 119.]])]
    heads = [synth.raw_logits(B, A, C, G, 70 + G) for G in grids]
    for h in heads:          # raise objectness so that a useful number of rows pass 0.25
        h.view(B, A, 5 + C, h.shape[2], h.shape[3])[:, :, 4] += 3.0
    dev_heads = [h.to(DEV) for h in heads]
    got, gidx = od.non_max_suppression(None, dev_heads, conf_thres=0.25, compat=False, decode="yolov5", anchors=anchors,
                                       strides=strides, return_index=True)
    rows = torch.cat([od.decode_box(h, a, s, "yolov5") for h, a, s in zip(dev_heads, anchors, strides)], 1).cpu()
    # class conf of the fused path is sigmoid(max logit) == max of the decoded sigmoids
    want, widx = rp.yolo_nms_fast(rows, conf_thres=0.25)
    for b in range(B):
        assert torch.equal(gidx[b].cpu(), widx[b])
        torch.testing.assert_close(got[b].cpu(), want[b], rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("G,C,off_out,off_head", [(20, 80, 1, 0), (20, 7, 2, 1), (13, 4, 3, 0), (6, 123, 1, 3), (7, 130, 1, 1)])
def test_decode_box_c_abi_unaligned_pointers_and_canaries(G, C, off_out, off_head):
    """The C entry takes any 4-byte aligned pointers: the tile kernel drops to 32-bit loads and matches its shared-memory
    tile to the destination's 16-byte phase.  Same values as the aligned call, nothing written outside [out, out + n)."""
    from objectdetectionpl_b200 import _lib as L
    lib = L.load()
    B, A, F = 2, 3, 5 + C
    head = synth.raw_logits(B, A, C, G, 5).to(DEV)
    anc = torch.tensor([[3.625, 2.8125], [4.875, 6.1875], [11.65625, 10.1875]], device=DEV)
    want = od.decode_box(head, anc, 16.0, "yolo_exp", num_anchors=A)
    n = want.numel()
    pad = 64
    buf = torch.full((n + 2 * pad,), 12345.0, device=DEV)
    hbuf = torch.zeros(head.numel() + 8, device=DEV)
    hbuf[off_head:off_head + head.numel()] = head.reshape(-1)
    out_ptr = buf.data_ptr() + 4 * (pad + off_out)
    rc = lib.b200det_decode_box(hbuf.data_ptr() + 4 * off_head, B, A, C, G, L.DECODE_YOLO_EXP, anc.data_ptr(), 16.0, out_ptr,
                                L.stream_ptr(head.device))
    assert rc == 0, lib.b200det_last_error()
    torch.cuda.synchronize()
    got = buf[pad + off_out: pad + off_out + n]
    assert torch.equal(got, want.reshape(-1))
    assert bool((buf[:pad + off_out] == 12345.0).all()) and bool((buf[pad + off_out + n:] == 12345.0).all())


def test_yolo_forward_dynamic_golden_d3():
    """D3: `yolo_forward_dynamic` of the unmodified reference (LightningFunc/utils/YoloV4Utils.py:36-176, scale_x_y 1.05)
    on the golden head; boxes [B,N,1,4] normalised corners, confs [B,N,C] = sigmoid(cls) * sigmoid(obj); 1e-5 relative."""
    d = load("decode")
    head = T(d["head"])
    A = 3
    C = head.shape[1] // A - 5
    flat = [float(v) for v in d["anchors"].reshape(-1)]
    boxes, confs = od.yolo_forward_dynamic(head.to(DEV), 0.5, C, flat, A, 1.05)
    assert tuple(boxes.shape) == tuple(d["d3_boxes"].shape) and tuple(confs.shape) == tuple(d["d3_confs"].shape)
    assert boxes.is_contiguous() and confs.is_contiguous()
    torch.testing.assert_close(boxes.cpu(), T(d["d3_boxes"]), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(confs.cpu(), T(d["d3_confs"]), rtol=1e-5, atol=1e-9)
    rb, rc = od.get_region_boxes([(boxes, confs), (boxes, confs)])
    assert tuple(rb.shape) == (boxes.shape[0], 2 * boxes.shape[1], 1, 4) and tuple(rc.shape) == (confs.shape[0], 2 * confs.shape[1], C)
    # rows form of the same kernel
    rows = od.decode_box(head.to(DEV), T(d["anchors"]), 1.0, "yolov4_norm", scale_x_y=1.05)
    assert torch.equal(rows[..., :4], boxes.squeeze(2)) and torch.equal(rows[..., 5:], confs)


@pytest.mark.parametrize("G,C,sxy", [(13, 4, 1.2), (20, 80, 1.05), (19, 7, 1.0), (8, 200, 1.1), (6, 1100, 1.05)])
def test_yolo_forward_dynamic_vs_oracle(G, C, sxy):
    B, A = 2, 3
    head = synth.raw_logits(B, A, C, G, 300 + G)
    anc = torch.tensor([[1.5, 2.0], [3.625, 2.8125], [4.875, 6.1875]])
    wb, wc = rp.decode_yolov4_norm(head, anc, scale_x_y=sxy)
    boxes, confs, det = od.yolo_forward_dynamic(head.to(DEV), 0.4, C, anc, A, sxy, return_det_confs=True)
    torch.testing.assert_close(boxes.cpu().squeeze(2), wb, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(confs.cpu(), wc, rtol=1e-5, atol=1e-9)
    want_det = torch.sigmoid(head.view(B, A, 5 + C, G * G)[:, :, 4]).reshape(B, -1)
    torch.testing.assert_close(det.cpu(), want_det, rtol=1e-5, atol=1e-9)


def test_fused_yolov4_norm_decode_nms_matches_decode_then_nms():
    """K1 mode DECODE_YOLOV4_NORM: non_max_suppression(decode='yolov4_norm') == oracle NMS on rows whose box is the D3
    corner box (as cx, cy, w, h), conf = sigmoid(obj), class scores = sigmoid(cls)."""
    B, A, C, G = 2, 3, 5, 16
    head = synth.raw_logits(B, A, C, G, 91)
    head.view(B, A, 5 + C, G, G)[:, :, 4] += 3.0
    anc = torch.tensor([[1.25, 1.625], [2.0, 3.75], [4.125, 2.875]])
    dev_head = head.to(DEV)
    got, gidx = od.non_max_suppression(None, [dev_head], conf_thres=0.25, compat=False, decode="yolov4_norm", anchors=[anc],
                                       return_index=True, scale_x_y=1.05)
    boxes, confs, det = od.yolo_forward_dynamic(dev_head, 0.4, C, anc, A, 1.05, return_det_confs=True)
    xyxy = boxes.squeeze(2).cpu()
    cls = torch.sigmoid(head.view(B, A, 5 + C, G * G)[:, :, 5:]).permute(0, 1, 3, 2).reshape(B, -1, C)
    # the oracle NMS wants cx, cy, w, h rows; the corner box of the fused path is (x1, y1, x1 + w, y1 + h)
    w, h = xyxy[..., 2] - xyxy[..., 0], xyxy[..., 3] - xyxy[..., 1]
    rows = torch.cat([torch.stack([xyxy[..., 0] + w / 2, xyxy[..., 1] + h / 2, w, h, det.cpu()], -1), cls], -1)
    want, widx = rp.yolo_nms_fast(rows, conf_thres=0.25)
    for b in range(B):
        assert torch.equal(gidx[b].cpu(), widx[b])
        assert torch.equal(got[b][:, 6].cpu(), want[b][:, 6])
        torch.testing.assert_close(got[b][:, :6].cpu(), want[b][:, :6], rtol=1e-4, atol=1e-5)

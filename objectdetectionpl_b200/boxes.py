"""Box maths with the reference's names and signatures (LightningFunc/accuracy.py), on the GPU.

  xywh2xyxy(x)                                           accuracy.py:289
  bbox_iou(box1, box2, x1y1x2y2=True, ...)               accuracy.py:39   (IoU with +1 pixel, +1e-16)
  iou(tens1, tens2)                                      accuracy.py:6    (plain corner IoU)
  bbox_iou_v5(box1, box2, x1y1x2y2, GIoU, DIoU, CIoU)    accuracy.py:71   (differentiable w.r.t. box1)
"""
from __future__ import annotations

import torch

from . import _lib as L


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    return L.require_cuda(t, name).contiguous()


def xywh2xyxy(x: torch.Tensor) -> torch.Tensor:
    x = _f32c(x, "x")
    if x.shape[-1] != 4:
        raise ValueError("last dimension must be 4")
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        L.check(L.load().b200det_xywh2xyxy(x.data_ptr(), y.data_ptr(), x.numel() // 4, L.stream_ptr(x.device)), "xywh2xyxy")
    return y


def bbox_iou(box1: torch.Tensor, box2: torch.Tensor, x1y1x2y2=True, GIoU=False, DIoU=False, CIoU=False) -> torch.Tensor:
    """IoU_+1 of box1 [1|n,4] against box2 [n,4].  GIoU/DIoU/CIoU are accepted and ignored, as in the
    reference (accuracy.py:39-69)."""
    box1, box2 = _f32c(box1, "box1"), _f32c(box2, "box2")
    if box1.dim() != 2 or box2.dim() != 2 or box1.shape[1] != 4 or box2.shape[1] != 4:
        raise ValueError("boxes must be [n,4]")
    n1, n = box1.shape[0], box2.shape[0]
    if n1 != 1 and n == 1:                      # broadcasting the other way round
        box2 = box2.expand(n1, 4).contiguous()
        n = n1
    out = torch.empty((n,), dtype=torch.float32, device=box2.device)
    with torch.cuda.device(box2.device):
        L.check(L.load().b200det_bbox_iou_plus1(box1.data_ptr(), n1, box2.data_ptr(), n, int(bool(x1y1x2y2)), out.data_ptr(),
                                                L.stream_ptr(box2.device)), "bbox_iou")
    return out


def iou(tens1: torch.Tensor, tens2: torch.Tensor) -> torch.Tensor:
    """Elementwise corner-box IoU on equally shaped [..., 4] tensors (accuracy.py:6-37)."""
    assert tens1.size() == tens2.size()
    assert tens1.size(-1) == 4
    a, b = _f32c(tens1, "tens1"), _f32c(tens2, "tens2")
    out = torch.empty(a.shape[:-1], dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        L.check(L.load().b200det_pair_iou(a.data_ptr(), b.data_ptr(), a.numel() // 4, out.data_ptr(), L.stream_ptr(a.device)),
                "iou")
    return out


class _BboxIouV5(torch.autograd.Function):
    @staticmethod
    def forward(ctx, box1, box2, corner, kind):
        b1 = L.require_cuda(box1.detach(), "box1")
        b2 = L.require_cuda(box2.detach(), "box2")
        if b1.dim() != 2 or b1.shape[0] != 4 or b2.shape != b1.shape:
            raise ValueError("bbox_iou_v5 expects transposed [4,n] boxes of equal shape")
        n = b1.shape[1]
        out = torch.empty((n,), dtype=torch.float32, device=b1.device)
        with torch.cuda.device(b1.device):
            L.check(L.load().b200det_bbox_iou_v5_fwd(b1.data_ptr(), b1.stride(0), b1.stride(1), b2.data_ptr(), b2.stride(0),
                                                     b2.stride(1), n, corner, kind, out.data_ptr(), L.stream_ptr(b1.device)),
                    "bbox_iou_v5")
        ctx.save_for_backward(b1, b2)
        ctx.corner, ctx.kind = corner, kind
        return out

    @staticmethod
    def backward(ctx, grad_out):
        b1, b2 = ctx.saved_tensors
        n = b1.shape[1]
        g = grad_out.contiguous().float()
        gin = torch.empty((4, n), dtype=torch.float32, device=b1.device)
        with torch.cuda.device(b1.device):
            L.check(L.load().b200det_bbox_iou_v5_bwd(b1.data_ptr(), b1.stride(0), b1.stride(1), b2.data_ptr(), b2.stride(0),
                                                     b2.stride(1), n, ctx.corner, ctx.kind, g.data_ptr(), gin.data_ptr(),
                                                     L.stream_ptr(b1.device)), "bbox_iou_v5 backward")
        return gin, None, None, None


def bbox_iou_v5(box1, box2, x1y1x2y2=True, GIoU=False, DIoU=False, CIoU=False):
    """IoU / GIoU / DIoU / CIoU of transposed `[4,n]` boxes (accuracy.py:71-114); `.t()` views are read in
    place.  Differentiable w.r.t. box1 (box2 is the target; CIoU's alpha is a constant, accuracy.py:110)."""
    kind = L.GIOU if GIoU else L.DIOU if DIoU else L.CIOU if CIoU else L.IOU      # precedence of accuracy.py:98-108
    return _BboxIouV5.apply(box1, box2, int(bool(x1y1x2y2)), kind)

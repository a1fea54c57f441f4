"""Loss-side target assignment with the reference's signatures, on the GPU.

  build_targets(pred_boxes, pred_cls, target, anchors, ignore_thres)   LightningFunc/accuracy.py:305
  build_targets_v5(p, targets, anchors, nl, na)                        LightningFunc/accuracy.py:472
  v5_match_level(pi, tbox, indices, anch)                              LightningFunc/losses.py:105-123
  v5_loss_level(...) / v5_loss(output, target, anchors, nl, na, nc)    LightningFunc/losses.py:98-152 (fused loss terms)
  ssd_match(default_boxes, annotations_boxes, match_thresh)            LightningFunc/losses.py:199
  retina_assign(anchors, targets, batch_size, img_size)                LightningFunc/losses.py:423-443
"""
from __future__ import annotations

import ctypes
import weakref

import torch

from . import _lib as L


# build_targets reads the kernel's status word (one 4-byte device->host copy, i.e. a stream sync) to raise the reference's
# IndexError on target rows whose negative indices cannot wrap into range.  Set to False to keep the call asynchronous:
# such rows then make every scatter be skipped (all-zero targets) without an exception.
CHECK_INDEX_STATUS = True


def build_targets(pred_boxes, pred_cls, target, anchors, ignore_thres):
    """Returns the reference's 10-tuple `(iou_scores, class_mask, obj_mask, noobj_mask, tx, ty, tw, th, tcls,
    tconf)` with identical dtypes (masks uint8).  Duplicate cells: the highest target row wins.  Index errors: see
    `CHECK_INDEX_STATUS`."""
    lib = L.load()
    pb = L.require_cuda(pred_boxes, "pred_boxes").contiguous()
    pc = L.require_cuda(pred_cls, "pred_cls").contiguous()
    tg = L.require_cuda(target, "target").contiguous()
    an = L.require_cuda(torch.as_tensor(anchors, dtype=torch.float32, device=pb.device), "anchors").contiguous()
    B, A, G = pb.shape[0], pb.shape[1], pb.shape[2]
    C, nt = pc.shape[-1], tg.shape[0]
    dev = pb.device
    f = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
    iou_scores, class_mask, tx, ty, tw, th = (f(B, A, G, G) for _ in range(6))
    tcls = f(B, A, G, G, C)
    obj = torch.empty((B, A, G, G), dtype=torch.uint8, device=dev)
    noobj = torch.empty((B, A, G, G), dtype=torch.uint8, device=dev)
    status = torch.empty((1,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        ws = L.workspace(lib.b200det_build_targets_workspace_bytes(B, A, G, nt), dev)
        L.check(lib.b200det_build_targets(pb.data_ptr(), pc.data_ptr(), tg.data_ptr() if nt else None, an.data_ptr(), B, A, G, C,
                                          nt, float(ignore_thres), ws.data_ptr(), ws.numel(), iou_scores.data_ptr(),
                                          class_mask.data_ptr(), obj.data_ptr(), noobj.data_ptr(), tx.data_ptr(), ty.data_ptr(),
                                          tw.data_ptr(), th.data_ptr(), tcls.data_ptr(), status.data_ptr(), L.stream_ptr(dev)),
                "build_targets")
    if CHECK_INDEX_STATUS and (int(status.item()) & 4):
        # a negative image / cell / label index below -size: the reference's advanced indexing raises here
        # (accuracy.py:345-374) instead of returning targets
        raise IndexError("build_targets: a target row indexes outside [-size, size) of the [B, A, G, G(, C)] maps")
    return iou_scores, class_mask, obj, noobj, tx, ty, tw, th, tcls, obj.float()


_ANCHOR_CACHE = {}          # id(tensor) -> (weakref to the tensor, its version, host array)


def _host_anchors(anchors, nl, na):
    """The criterion's scaled anchors as a host float array for the launch parameters.  They are constants of the model:
    a device tensor is copied to the host once per tensor object and version (the copy is a host sync), not every step."""
    if isinstance(anchors, torch.Tensor) and anchors.is_cuda:
        hit = _ANCHOR_CACHE.get(id(anchors))
        if hit is not None and hit[0]() is anchors and hit[1] == anchors._version:
            return hit[2]
        if len(_ANCHOR_CACHE) > 64:
            _ANCHOR_CACHE.clear()
        vals = anchors.detach().float().cpu().reshape(nl, na, 2).reshape(-1).tolist()
        arr = (ctypes.c_float * (2 * na * nl))(*vals)
        _ANCHOR_CACHE[id(anchors)] = (weakref.ref(anchors), anchors._version, arr)
        return arr
    vals = torch.as_tensor(anchors).detach().float().reshape(nl, na, 2).reshape(-1).tolist()
    return (ctypes.c_float * (2 * na * nl))(*vals)


def _build_targets_v5_raw(p, targets, anchors, nl, na, sync=True):
    """One launch for all levels, one host sync for the row counts.  Per level: `(ib int32 [5, m] = b, a, gj, gi, cls
    (rows of a wider buffer), tbox [m, 4], anch [m, 2])`.  `sync=False`: no host sync — returns the full-capacity buffers
    `[(ib [5, cap], tbox [cap, 4], anch [cap, 2])]` and the device tensor of the per-level row counts."""
    lib = L.load()
    tg = L.require_cuda(targets, "targets").contiguous()
    dev = tg.device
    nt = tg.shape[0]
    arr = _host_anchors(anchors, nl, na)
    cap = max(5 * na * nt, 1)
    counts = torch.empty((nl,), dtype=torch.int32, device=dev)
    nxs, nys = (ctypes.c_int32 * nl)(), (ctypes.c_int32 * nl)()
    ptrs = [(ctypes.c_void_p * nl)() for _ in range(7)]                  # b, a, gj, gi, cls, tbox, anch per level
    # one buffer per kind for all levels (three allocations per call, addresses by arithmetic: no per-level tensor ops)
    ib_all = torch.empty((nl, 5, cap), dtype=torch.int32, device=dev)    # b, a, gj, gi, cls
    tb_all = torch.empty((nl, cap, 4), dtype=torch.float32, device=dev)
    ac_all = torch.empty((nl, cap, 2), dtype=torch.float32, device=dev)
    ib0, tb0, ac0 = ib_all.data_ptr(), tb_all.data_ptr(), ac_all.data_ptr()
    for i in range(nl):
        shape = p[i].shape if isinstance(p[i], torch.Tensor) else tuple(p[i])
        nys[i], nxs[i] = int(shape[2]), int(shape[3])
        for k in range(5):
            ptrs[k][i] = ib0 + 4 * cap * (5 * i + k)
        ptrs[5][i], ptrs[6][i] = tb0 + 16 * cap * i, ac0 + 8 * cap * i
    if torch.cuda.current_device() != dev.index:
        with torch.cuda.device(dev):
            L.check(lib.b200det_build_targets_v5(tg.data_ptr() if nt else None, nt, nl, arr, na, nxs, nys, *ptrs,
                                                 counts.data_ptr(), L.stream_ptr(dev)), "build_targets_v5")
    else:
        L.check(lib.b200det_build_targets_v5(tg.data_ptr() if nt else None, nt, nl, arr, na, nxs, nys, *ptrs, counts.data_ptr(),
                                             L.stream_ptr(dev)), "build_targets_v5")     # one launch, a cluster of 8 CTAs per level
    if not sync:
        return _V5Rows(ib_all, tb_all, ac_all, cap, nl), counts
    ms = counts.cpu().tolist()                                            # one host sync for all levels
    return [(ib_all[i, :, :m], tb_all[i, :m], ac_all[i, :m]) for i, m in enumerate(ms)]


class _V5Rows:
    """The full-capacity target rows of all levels (`sync=False` form of `_build_targets_v5_raw`): three buffers and the
    addresses of every level's arrays in them.  `ptrs(i)` = (b, a, gj, gi, cls, tbox, anch) of level i."""
    __slots__ = ("ib", "tb", "ac", "cap", "nl", "_ptrs")

    def __init__(self, ib, tb, ac, cap, nl):
        self.ib, self.tb, self.ac, self.cap, self.nl = ib, tb, ac, cap, nl
        ib0, tb0, ac0 = ib.data_ptr(), tb.data_ptr(), ac.data_ptr()
        self._ptrs = [tuple(ib0 + 4 * cap * (5 * i + k) for k in range(5)) + (tb0 + 16 * cap * i, ac0 + 8 * cap * i)
                      for i in range(nl)]

    def ptrs(self, i):
        return self._ptrs[i]


def build_targets_v5(p, targets, anchors, nl, na):
    """Returns `(tcls, tbox, indices, anch)`, four lists of length nl with the reference's dtypes
    (int64 indices/classes) and row order.  `p[i]` may be a tensor `[B,na,ny,nx,5+C]` or just its shape."""
    tcls, tbox, indices, anch = [], [], [], []
    for ib, tb, ac in _build_targets_v5_raw(p, targets, anchors, nl, na):
        il = ib.long()
        indices.append((il[0], il[1], il[2], il[3]))
        tcls.append(il[4])
        tbox.append(tb)
        anch.append(ac)
    return tcls, tbox, indices, anch


class _V5Match(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pi, tbox, b, a, gj, gi, anch):
        lib = L.load()
        pid = L.require_cuda(pi.detach(), "pi")
        if not pid.is_contiguous():
            raise ValueError("pi must be contiguous [B,na,ny,nx,5+C]")
        B, na, ny, nx, F = pid.shape
        m = b.shape[0]
        dev = pid.device
        idx = torch.stack((b, a, gj, gi)).to(torch.int32).contiguous()
        tb = tbox.detach().contiguous().float()
        ac = anch.detach().contiguous().float()
        giou = torch.empty((m,), dtype=torch.float32, device=dev)
        tobj = torch.zeros((B, na, ny, nx), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            L.check(lib.b200det_v5_match_fwd(pid.data_ptr(), B, na, ny, nx, F, idx[0].data_ptr(), idx[1].data_ptr(),
                                             idx[2].data_ptr(), idx[3].data_ptr(), tb.data_ptr(), ac.data_ptr(), m, giou.data_ptr(),
                                             tobj.data_ptr(), L.stream_ptr(dev)), "v5_match_fwd")
        ctx.save_for_backward(pid, idx, tb, ac)
        ctx.mark_non_differentiable(tobj)
        return giou, tobj

    @staticmethod
    def backward(ctx, g_giou, _g_tobj):
        lib = L.load()
        pid, idx, tb, ac = ctx.saved_tensors
        B, na, ny, nx, F = pid.shape
        m = idx.shape[1]
        gpi = torch.zeros_like(pid)
        g = g_giou.contiguous().float()
        with torch.cuda.device(pid.device):
            L.check(lib.b200det_v5_match_bwd(pid.data_ptr(), B, na, ny, nx, F, idx[0].data_ptr(), idx[1].data_ptr(),
                                             idx[2].data_ptr(), idx[3].data_ptr(), tb.data_ptr(), ac.data_ptr(), m, g.data_ptr(),
                                             gpi.data_ptr(), L.stream_ptr(pid.device)), "v5_match_bwd")
        return gpi, None, None, None, None, None, None


def v5_match_level(pi, tbox, indices, anch):
    """Fused matched-row path of `MultiScaleRegionLoss_v5.forward` for one level (losses.py:105-123):
    gather `pi[b,a,gj,gi]`, decode (σ·2−0.5, (σ·2)²·anchor), GIoU against `tbox`, and the objectness target
    `tobj[b,a,gj,gi] = clamp(giou,0)` (last row wins on duplicate cells).  Returns `(giou[m], tobj[B,na,ny,nx])`;
    `giou` is differentiable w.r.t. `pi`."""
    b, a, gj, gi = indices
    return _V5Match.apply(pi, tbox, b, a, gj, gi, anch)


class _V5LossLevel(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pi, tbox, idx, anch, cp, cn, gamma, alpha, with_cls):
        # idx: int32 [5, m] = b, a, gj, gi, cls; each row contiguous (rows of a wider buffer are fine)
        lib = L.load()
        pid = L.require_cuda(pi.detach(), "pi")
        if not pid.is_contiguous():
            raise ValueError("pi must be contiguous [B,na,ny,nx,5+C]")
        B, na, ny, nx, F = pid.shape
        m = int(idx.shape[1])
        dev = pid.device
        tb = tbox.detach().contiguous().float()
        ac = anch.detach().contiguous().float()
        giou = torch.empty((max(m, 1),), dtype=torch.float32, device=dev)
        tobj = torch.empty((B, na, ny, nx), dtype=torch.float32, device=dev)
        sums = torch.empty((3,), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            L.check(lib.b200det_v5_loss_fwd(pid.data_ptr(), B, na, ny, nx, F, idx[0].data_ptr(), idx[1].data_ptr(),
                                            idx[2].data_ptr(), idx[3].data_ptr(), idx[4].data_ptr(), tb.data_ptr(), ac.data_ptr(),
                                            m, cp, cn, gamma, alpha, int(with_cls), giou.data_ptr(), tobj.data_ptr(),
                                            sums.data_ptr(), L.stream_ptr(dev)), "v5_loss_fwd")
        cells = B * na * ny * nx
        n_box, n_cls = max(m, 1), max(m * (F - 5), 1)
        means = sums.float()                                              # the launcher's last kernel divided by the counts
        ctx.save_for_backward(pid, idx, tb, ac, tobj)
        ctx.cfg = (cp, cn, gamma, alpha, int(with_cls), m, cells, n_box, n_cls)
        ctx.mark_non_differentiable(tobj)
        return means[0], means[1], means[2], tobj

    @staticmethod
    def backward(ctx, g_box, g_obj, g_cls, _g_tobj):
        lib = L.load()
        pid, idx, tb, ac, tobj = ctx.saved_tensors
        cp, cn, gamma, alpha, with_cls, m, cells, n_box, n_cls = ctx.cfg
        B, na, ny, nx, F = pid.shape
        gpi = torch.empty_like(pid)                                       # fully written by the call (no zero-fill pass)
        g3 = torch.stack((g_box, g_obj, g_cls)).float().contiguous()      # stays on the device: no sync in backward
        with torch.cuda.device(pid.device):
            L.check(lib.b200det_v5_loss_bwd_full(pid.data_ptr(), B, na, ny, nx, F, idx[0].data_ptr(), idx[1].data_ptr(),
                                            idx[2].data_ptr(), idx[3].data_ptr(), idx[4].data_ptr(), tb.data_ptr(), ac.data_ptr(),
                                            m, cp, cn, gamma, alpha, with_cls, tobj.data_ptr(), g3.data_ptr(), 1.0 / n_box,
                                            1.0 / cells, 1.0 / n_cls, gpi.data_ptr(), L.stream_ptr(pid.device)), "v5_loss_bwd")
        return (gpi,) + (None,) * 8


def v5_loss_level(pi, tbox, indices, anch, tcls, cp=1.0, cn=0.0, gamma=1.5, alpha=0.25, with_cls=True):
    """The three loss terms of one level of `MultiScaleRegionLoss_v5.forward` (losses.py:105-137), fused and
    differentiable w.r.t. `pi`: returns `(mean(1 - giou), mean FL(pi[...,4], tobj), mean FL(ps[:,5:], class targets), tobj)`.
    With no matched rows the first and third are 0 (the reference skips them, :110)."""
    b, a, gj, gi = indices
    if b.shape[0]:
        idx = torch.stack((b, a, gj, gi, tcls)).to(torch.int32)
    else:
        idx = torch.zeros((5, 0), dtype=torch.int32, device=pi.device)
    return _V5LossLevel.apply(pi, tbox, idx, anch, float(cp), float(cn), float(gamma), float(alpha), bool(with_cls))


_V5_GAINS = (0.05, 1.0, 0.58)          # box, obj, cls  (losses.py:139-141)


class _V5LossAll(torch.autograd.Function):
    """All levels and the combination of `MultiScaleRegionLoss_v5.forward` as ONE autograd node: the per-level Python and
    the ~15 tiny autograd nodes of the level-by-level form were two thirds of its wall time.  The matched-row counts stay on
    the device (`b200det_v5_loss_fwd_dev` / `_bwd_full_dev`): nothing between `build_targets_v5` and the loss value waits for
    the GPU."""

    @staticmethod
    def forward(ctx, levels, counts, cfg, *pis):
        lib = L.load()
        cp, cn, gamma, alpha, with_cls = cfg
        nl = len(pis)
        pids = []
        for pi in pis:
            pid = L.require_cuda(pi.detach(), "pi")
            if not pid.is_contiguous():
                raise ValueError("pi must be contiguous [B,na,ny,nx,5+C]")
            pids.append(pid)
        dev = pids[0].device
        if torch.cuda.current_device() != dev.index:
            with torch.cuda.device(dev):
                return _V5LossAll.forward(ctx, levels, counts, cfg, *pis)
        ctx.set_materialize_grads(False)            # the three metric outputs usually get no gradient: None, not a zero tensor each
        cells = [pid.numel() // pid.shape[-1] for pid in pids]
        cap = levels.cap
        lv = (L.V5Level * nl)()
        for i, pid in enumerate(pids):
            l = lv[i]
            l.pi = pid.data_ptr()
            l.batch, l.na, l.ny, l.nx, l.fields = pid.shape
            l.b, l.a, l.gj, l.gi, l.tcls, l.tbox, l.anch = levels.ptrs(i)
            l.m_dev = counts.data_ptr() + 4 * i
        tobj = torch.empty((sum(cells),), dtype=torch.float32, device=dev)            # all levels, back to back (scratch)
        gobj = torch.empty((sum(cells),), dtype=torch.float32, device=dev)            # d FL_obj / d logit per cell, for the backward
        giou = torch.empty((nl * cap,), dtype=torch.float32, device=dev)
        means = torch.empty((nl, 3), dtype=torch.float64, device=dev)
        out = torch.empty((4,), dtype=torch.float32, device=dev)
        L.check(lib.b200det_v5_loss_fwd_all(lv, nl, cap, cp, cn, gamma, alpha, int(with_cls), *_V5_GAINS, giou.data_ptr(),
                                            tobj.data_ptr(), gobj.data_ptr(), means.data_ptr(), out.data_ptr(), L.stream_ptr(dev)),
                "v5_loss_fwd_all")                 # 2 memsets + 5 launches for all levels, the combination included
        ctx.save_for_backward(gobj, counts, *pids)         # tobj is not needed again: the backward gets its effect through gobj
        ctx.levels, ctx.cfg, ctx.lv = levels, cfg, lv
        return out[0:1], out[1:2], out[2:3], out[3:4]

    @staticmethod
    def backward(ctx, g_loss, g_box, g_cls, g_obj):
        lib = L.load()
        gobj, counts, *pids = ctx.saved_tensors
        cp, cn, gamma, alpha, with_cls = ctx.cfg
        dev = gobj.device
        if torch.cuda.current_device() != dev.index:
            with torch.cuda.device(dev):
                return _V5LossAll.backward(ctx, g_loss, g_box, g_cls, g_obj)
        gs = [None if g is None else g.contiguous().float() for g in (g_loss, g_box, g_cls, g_obj)]
        g3 = torch.empty((3,), dtype=torch.float32, device=dev)
        lv, nl = ctx.lv, len(pids)
        grads = [torch.empty_like(pid) for pid in pids]               # fully written by the call (no zero-fill pass)
        for i, gpi in enumerate(grads):
            lv[i].gpi = gpi.data_ptr()
        L.check(lib.b200det_v5_loss_bwd_all(lv, nl, ctx.levels.cap, cp, cn, gamma, alpha, int(with_cls), *_V5_GAINS, gobj.data_ptr(),
                                            *(None if g is None else g.data_ptr() for g in gs), g3.data_ptr(), L.stream_ptr(dev)),
                "v5_loss_bwd_all")                 # 3 launches for all levels
        return (None, None, None, *grads)


def v5_loss(output, target, anchors, nl, na, nc, cp=1.0, cn=0.0, gamma=1.5, alpha=0.25):
    """`MultiScaleRegionLoss_v5.forward` (losses.py:98-152, reduction 'mean', label smoothing 0, focal gamma 1.5):
    `build_targets_v5` (one launch), the fused per-level loss kernels and the gain-weighted combination, as one autograd
    node and without a host sync (the row counts stay on the device).  `anchors` are the criterion's scaled anchors
    `[nl, na, 2]` (:95-96).  Returns the reference's metrics dict of shape-[1] tensors: loss, Localization, Classification,
    Conf_obj."""
    levels, counts = _build_targets_v5_raw(output, target, anchors, nl, na, sync=False)   # full-capacity rows + device counts
    cfg = (float(cp), float(cn), float(gamma), float(alpha), nc > 1)
    loss, lbox, lcls, lobj = _V5LossAll.apply(levels, counts, cfg, *output)
    return {"loss": loss, "Localization": lbox, "Classification": lcls, "Conf_obj": lobj}


def ssd_match(default_boxes, annotations_boxes, match_thresh=0.5):
    """`SSDLoss.match` (losses.py:199-218): returns `(box_with_annotation[P] int64, matched[P] bool)`."""
    lib = L.load()
    pri = L.require_cuda(default_boxes, "default_boxes").contiguous()
    gt = L.require_cuda(annotations_boxes, "annotations_boxes").contiguous()
    P, M = pri.shape[0], gt.shape[0]
    dev = pri.device
    idx = torch.empty((P,), dtype=torch.int32, device=dev)
    matched = torch.empty((P,), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        ws = L.workspace(lib.b200det_ssd_match_workspace_bytes(P, M), dev)
        L.check(lib.b200det_ssd_match(pri.data_ptr(), P, gt.data_ptr() if M else None, M, float(match_thresh), ws.data_ptr(),
                                      ws.numel(), idx.data_ptr(), matched.data_ptr(), L.stream_ptr(dev)), "ssd_match")
    return idx.long(), matched.bool()


def retina_assign(anchors, targets, batch_size, img_size):
    """RetinaNet target assignment + encoding (losses.py:423-443): returns
    `(loc_targets [B,A,4] fp32, cls_targets [B,A] int64)`; cls = 1+label, 0 background (<0.5), −1 ignore (0.4..0.5)."""
    lib = L.load()
    an = L.require_cuda(anchors, "anchors").contiguous()
    tg = L.require_cuda(targets, "targets").contiguous()
    A, nt, dev = an.shape[0], tg.shape[0], an.device
    loc = torch.empty((batch_size, A, 4), dtype=torch.float32, device=dev)
    cls = torch.empty((batch_size, A), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        ws = L.workspace(lib.b200det_retina_assign_workspace_bytes(batch_size, nt), dev)
        L.check(lib.b200det_retina_assign(an.data_ptr(), A, tg.data_ptr() if nt else None, nt, batch_size, float(img_size),
                                          ws.data_ptr(), ws.numel(), loc.data_ptr(), cls.data_ptr(), L.stream_ptr(dev)),
                "retina_assign")
    return loc, cls.long()

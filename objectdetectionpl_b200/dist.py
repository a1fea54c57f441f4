"""Multi-GPU glue: the batch is sharded BY IMAGE (every image is independent in decode, filter, sort and
NMS — SURVEY.md §8e), so the hot path needs no collective.  The only exchange step is after it: an
all-gather of the detections for mAP evaluation (LightningFunc/step.py:95,102-130 consume the detections of all
images together).

The exchange is two collectives and one host read:
  1. `all_gather_into_tensor` of the per-image counts                [W, B_max] int32
  2. ONE device->host copy of those W*B_max integers (sizes the payload and gives every segment offset)
  3. `all_gather_into_tensor` of the packed rows `[K_max, 8]`       (7 detection columns + global image id) into a
     persistent `[W, K_max, 8]` buffer; the rows are packed on the device by `b200det_pack_detections` straight from the
     padded NMS output, no Python loop over images.
`GatheredDetections.per_image()` is then pure slicing by the known offsets.  Works with any torch.distributed backend
(NCCL over NVLink on the B200 box; gloo in the CPU tests, where the rows are packed with torch ops)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _lib as L


def shard_range(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous image block [lo, hi) of `rank`; the first `batch % world` ranks get one extra image."""
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_targets(targets: torch.Tensor, lo: int, hi: int) -> torch.Tensor:
    """Rows of `targets[nt,6]` whose image id is in [lo, hi), re-based so the shard's first image is 0."""
    sel = (targets[:, 0] >= lo) & (targets[:, 0] < hi)
    out = targets[sel].clone()
    out[:, 0] -= lo
    return out


def pack_detections(dets: Sequence[Optional[torch.Tensor]], image_offset: int, device=None) -> torch.Tensor:
    """List of per-image `[K,7]` rows (or None) -> one `[sum K, 8]` tensor, col 7 = global image id."""
    packed, _ = _pack_list(dets, image_offset, device)
    return packed


def _pack_list(dets, image_offset, device=None):
    """-> (packed [K,8], per-image counts int32 [B]) with two torch ops (one cat, one repeat_interleave)."""
    dev = device if device is not None else next((d.device for d in dets if d is not None), torch.device("cpu"))
    counts = torch.tensor([0 if d is None else int(d.shape[0]) for d in dets], dtype=torch.int32)
    rows = [d for d in dets if d is not None and d.shape[0]]
    if not rows:
        return torch.zeros((0, 8), dtype=torch.float32, device=dev), counts.to(dev)
    body = torch.cat(rows, 0)
    ids = torch.repeat_interleave(torch.arange(len(dets), dtype=torch.float32) + float(image_offset), counts.long())
    return torch.cat([body, ids.to(body.device).unsqueeze(1)], 1), counts.to(body.device)


def pack_detections_raw(rows: torch.Tensor, count: torch.Tensor, image_offset: int, out: Optional[torch.Tensor] = None,
                        max_rows: Optional[int] = None):
    """Device-side packing of the padded NMS output (`yolo_nms_raw` / `prior_nms_raw`: rows `[B, pitch, 7]`, count `[B]`
    int32) into `out [cap, 8]` (allocated for the worst case `B * pitch` if not given).  Returns `(out, offsets [B+1] int32)`,
    all on the device, nothing synchronised."""
    lib = L.load()
    L.require_cuda(rows, "rows")
    L.require_cuda(count, "count", torch.int32)
    if rows.dim() != 3 or rows.shape[2] != 7 or not rows.is_contiguous():
        raise ValueError(f"rows must be contiguous [B, pitch, 7], got {tuple(rows.shape)}")
    B, pitch = rows.shape[0], rows.shape[1]
    dev = rows.device
    if out is None:
        out = torch.empty((B * pitch, 8), dtype=torch.float32, device=dev)
    offsets = torch.empty((B + 1,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.b200det_pack_detections(rows.data_ptr(), count.data_ptr(), B, pitch,
                                            int(pitch if max_rows is None else min(max_rows, pitch)), int(image_offset),
                                            out.data_ptr(), out.shape[0], offsets.data_ptr(), L.stream_ptr(dev)),
                "pack_detections")
    return out, offsets


class GatheredDetections:
    """Result of the exchange step: `rows [W, K_max, 8]` (rank r's detections in rows[r, :totals[r]], images in order,
    score order inside an image; column 7 = global image id) and `counts [W, B_max]` on the HOST."""

    def __init__(self, rows: torch.Tensor, counts: torch.Tensor, batches: Sequence[int]):
        self.rows, self.counts, self.batches = rows, counts, list(batches)
        self.totals = [int(counts[r, :b].sum()) for r, b in enumerate(self.batches)]

    def per_image(self) -> List[Optional[torch.Tensor]]:
        """One entry per global image (rank order, then the rank's image order): `[K_i, 7]` view or None — slices by the
        offsets the counts give, no masking."""
        out: List[Optional[torch.Tensor]] = []
        for r, b in enumerate(self.batches):
            off = 0
            for k in self.counts[r, :b].tolist():
                out.append(self.rows[r, off:off + k, :7] if k else None)
                off += k
        return out

    def packed(self) -> torch.Tensor:
        """`[K_total, 8]`: the padding between the ranks' blocks squeezed out (one cat of W slices)."""
        return torch.cat([self.rows[r, :t] for r, t in enumerate(self.totals)], 0)


_gather_bufs = {}     # (device, group id) -> persistent flat receive buffer


def _recv_buffer(dev, group, numel):
    key = (str(dev), id(group))
    buf = _gather_bufs.get(key)
    if buf is None or buf.numel() < numel:
        buf = torch.empty(int(numel * 1.25) + 1024, dtype=torch.float32, device=dev)
        _gather_bufs[key] = buf
    return buf


def _all_gather_into(out: torch.Tensor, inp: torch.Tensor, group):
    try:
        dist.all_gather_into_tensor(out, inp, group=group)
    except (RuntimeError, NotImplementedError):      # a backend without the flat form (old gloo)
        parts = list(out.unbind(0))
        dist.all_gather(parts, inp, group=group)


def exchange(local: torch.Tensor, counts: torch.Tensor, group=None, batch_max: Optional[int] = None) -> GatheredDetections:
    """`local [>=K,8]` packed rows of this rank (capacity may exceed K), `counts [B]` int32 per-image counts (device or
    host).  All ranks get everybody's rows.  Uneven shards: pass `batch_max` = the largest shard size (default: same B on
    every rank, checked)."""
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    dev = local.device
    B = counts.shape[0]
    Bm = int(batch_max) if batch_max is not None else B
    mine = torch.zeros((Bm + 1,), dtype=torch.int32, device=dev)
    mine[:B] = counts.to(dev)
    mine[Bm] = B
    if world == 1:
        allc = mine.unsqueeze(0)
    else:
        allc = torch.empty((world, Bm + 1), dtype=torch.int32, device=dev)
        _all_gather_into(allc, mine, group)
    host = allc.cpu()                                           # the one host read of the exchange
    batches = host[:, Bm].tolist()
    cnt = host[:, :Bm].long()
    totals = [int(cnt[r, :b].sum()) for r, b in enumerate(batches)]
    kmax = max(max(totals), 1)
    if world == 1:
        return GatheredDetections(local[:kmax].unsqueeze(0), cnt, batches)
    if local.shape[0] < kmax:                                   # this rank holds fewer rows than the largest: pad its send buffer
        pad = torch.zeros((kmax, 8), dtype=torch.float32, device=dev)
        pad[:local.shape[0]] = local
        local = pad
    recv = _recv_buffer(dev, group, world * kmax * 8)[: world * kmax * 8].view(world, kmax, 8)
    _all_gather_into(recv, local[:kmax].contiguous(), group)
    return GatheredDetections(recv, cnt, batches)


def gather_detections_raw(rows: torch.Tensor, count: torch.Tensor, image_offset: int, group=None,
                          batch_max: Optional[int] = None, send: Optional[torch.Tensor] = None) -> GatheredDetections:
    """The exchange step on the padded NMS output (device tensors of `yolo_nms_raw`): device-side pack, counts gather, one
    host read, rows gather.  `send` is an optional persistent `[cap, 8]` pack buffer."""
    packed, _ = pack_detections_raw(rows, count, image_offset, out=send)
    return exchange(packed, count, group, batch_max)


def gather_detections(dets: Sequence[Optional[torch.Tensor]], image_offset: int, group=None, device=None,
                      batch_max: Optional[int] = None) -> torch.Tensor:
    """All-gather every rank's per-image detection list: returns `[K_total, 8]` (rank order, then image, then score
    order) on every rank.  (List form of the reference's `suppress_output`; `gather_detections_raw` skips the list.)"""
    local, counts = _pack_list(dets, image_offset, device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    if batch_max is None:                                       # uneven shards: agree on the largest shard first
        b = torch.tensor([len(dets)], dtype=torch.int64, device=local.device)
        dist.all_reduce(b, op=dist.ReduceOp.MAX, group=group)
        batch_max = int(b.item())
    return exchange(local, counts, group, batch_max).packed()


def unpack_detections(packed: torch.Tensor, batch: int) -> List[Optional[torch.Tensor]]:
    """Inverse of pack/gather: `[K,8]` (rows grouped by ascending image id, as every producer here writes them) -> list of
    length `batch` of `[K_i,7]` views (or None).  One bincount + one host read, then slicing."""
    out: List[Optional[torch.Tensor]] = [None] * batch
    if packed.shape[0] == 0:
        return out
    img = packed[:, 7].long()
    info = torch.cat([torch.bincount(img, minlength=batch), (img[1:] < img[:-1]).any().long().reshape(1)]).cpu().tolist()
    counts = info[:-1]
    if info[-1]:                                                # not grouped by image: order first (stable keeps the score order)
        packed = packed[torch.argsort(img, stable=True)]
    off = 0
    for i, k in enumerate(counts[:batch]):
        if k:
            out[i] = packed[off:off + k, :7]
        off += k
    return out

"""Multi-GPU glue: the batch is sharded BY IMAGE (every image is independent in decode, filter, sort and
NMS — SURVEY.md §8e), so the hot path needs no collective.  The only exchange step is after it: an
all-gather of the detections for mAP evaluation (counts, then rows padded to the global maximum).
Works with any torch.distributed backend (NCCL over NVLink on the B200 box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


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
    rows = []
    for i, d in enumerate(dets):
        if d is None or d.shape[0] == 0:
            continue
        rows.append(torch.cat([d, torch.full((d.shape[0], 1), float(image_offset + i), dtype=d.dtype, device=d.device)], 1))
    if rows:
        return torch.cat(rows, 0)
    dev = device if device is not None else next((d.device for d in dets if d is not None), torch.device("cpu"))
    return torch.zeros((0, 8), dtype=torch.float32, device=dev)


def gather_detections(dets: Sequence[Optional[torch.Tensor]], image_offset: int, group=None, device=None) -> torch.Tensor:
    """All-gather every rank's detections: returns `[K_total, 8]` (rank order, then image, then score order)
    on every rank.  Two collectives: counts `[W]`, then rows padded to the global max."""
    local = pack_detections(dets, image_offset, device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    kmax = max(max(counts), 1)
    padded = torch.zeros((kmax, 8), dtype=torch.float32, device=local.device)
    padded[:local.shape[0]] = local
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded, group=group)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], 0)


def unpack_detections(packed: torch.Tensor, batch: int) -> List[Optional[torch.Tensor]]:
    """Inverse of pack/gather: `[K,8]` -> list of length `batch` of `[K_i,7]` (or None)."""
    out: List[Optional[torch.Tensor]] = [None] * batch
    if packed.shape[0] == 0:
        return out
    img = packed[:, 7].long()
    for i in torch.unique(img).tolist():
        out[i] = packed[img == i, :7]
    return out

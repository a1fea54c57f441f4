"""CPU, world_size 2 (gloo): the image-sharded multi-GPU path.  Every rank post-processes its own block of
images (here with the CPU oracle standing in for the CUDA kernels, which is allowed in tests) and the
detections are all-gathered for mAP; the result must equal the single-process result on the whole batch."""
import os
import socket
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from objectdetectionpl_b200 import dist as odist, synth
from oracle import ref_port as rp

B, A, C, GRIDS, IMG = 5, 3, 4, [8, 4], 64      # 5 images over 2 ranks -> uneven shards (3 + 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        levels = synth.yolo_planar(B, A, C, GRIDS, IMG, seed=99)
        lo, hi = odist.shard_range(B, rank, world)
        shard = [t[lo:hi] for t in levels]
        dets = rp.yolo_nms(shard, num_anchors=A)
        gathered = odist.gather_detections(dets, image_offset=lo)
        local, counts = odist._pack_list(dets, lo)
        g = odist.exchange(local, counts, batch_max=3)           # the device path's host logic: counts -> offsets -> slices
        per_image = [None if t is None else t.clone() for t in g.per_image()]
        tg = synth.labels(B, C, seed=5, max_per_image=4)
        local_t = odist.shard_targets(tg, lo, hi)
        counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([local_t.shape[0]]))
        torch.save({"gathered": gathered, "nt": [int(c) for c in counts], "lo": lo, "hi": hi, "per_image": per_image,
                    "batches": g.batches, "totals": g.totals, "local_t": local_t}, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    for batch in (1, 5, 64, 511, 512):
        for world in (1, 2, 3, 8):
            spans = [odist.shard_range(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert odist.shard_range(512, 3, 8) == (192, 256)


def test_pack_unpack_roundtrip():
    dets = [torch.rand(3, 7), None, torch.rand(1, 7), torch.zeros(0, 7)]
    packed = odist.pack_detections(dets, image_offset=10)
    assert packed.shape == (4, 8) and packed[:, 7].tolist() == [10.0, 10.0, 10.0, 12.0]
    back = odist.unpack_detections(packed, 14)
    assert torch.equal(back[10], dets[0]) and torch.equal(back[12], dets[2]) and back[11] is None and back[13] is None
    assert odist.pack_detections([None, None], 0).shape == (0, 8)
    # rows that are not grouped by image are ordered first (stable: the score order inside an image survives)
    shuffled = torch.cat([packed[3:], packed[:3]])
    back2 = odist.unpack_detections(shuffled, 14)
    assert torch.equal(back2[10], dets[0]) and torch.equal(back2[12], dets[2])


def test_two_rank_gather_equals_single_process():
    world = 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), d), nprocs=world, join=True)
        res = [torch.load(os.path.join(d, f"rank{r}.pt")) for r in range(world)]
    levels = synth.yolo_planar(B, A, C, GRIDS, IMG, seed=99)
    whole = rp.yolo_nms(levels, num_anchors=A)
    want = odist.pack_detections(whole, image_offset=0)
    for r in res:                                   # identical on every rank, and equal to the unsharded result
        assert torch.equal(r["gathered"], want)
    assert torch.equal(res[0]["gathered"], res[1]["gathered"])
    back = odist.unpack_detections(res[0]["gathered"], B)
    for i in range(B):
        assert torch.equal(back[i], whole[i])
    for r in res:                                   # GatheredDetections.per_image(): offset slicing, uneven shards (3 + 2)
        assert r["batches"] == [3, 2] and sum(r["totals"]) == want.shape[0] and len(r["per_image"]) == B
        for i in range(B):
            assert torch.equal(r["per_image"][i], whole[i])
    tg = synth.labels(B, C, seed=5, max_per_image=4)
    assert sum(res[0]["nt"]) == tg.shape[0]
    for r in res:                                   # targets re-based to the shard
        assert float(r["local_t"][:, 0].min()) >= 0 and float(r["local_t"][:, 0].max()) < r["hi"] - r["lo"]

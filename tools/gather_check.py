"""torchrun script (one rank per GPU): the N-way sharded dense-crowd run followed by the detection all-gather must equal the
single-GPU run on the whole batch, bit for bit (BASELINE config 5's exchange step; LightningFunc/step.py:95,102-130 consume
the gathered set).  Prints one JSON line on rank 0.  Used by tests/test_gpu_multi.py and by hand:
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 tools/gather_check.py --per-rank 4 --img 1280"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import objectdetectionpl_b200 as od
from objectdetectionpl_b200 import synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--per-rank", type=int, default=2)
    ap.add_argument("--img", type=int, default=640)
    ap.add_argument("--classes", type=int, default=5)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    B = a.per_rank * world
    grids = [a.img // 8, a.img // 16, a.img // 32]
    whole = synth.yolo_crowd(B, 3, a.classes, grids, a.img, seed=5)            # same bytes on every rank
    lo, hi = od.dist.shard_range(B, rank, world)
    shard = [t[lo:hi].contiguous().to(dev) for t in whole]
    rows, _, count = od.yolo_nms_raw(shard, 3, conf_thres=0.001)
    g = od.dist.gather_detections_raw(rows, count, lo)
    per = g.per_image()
    ok, detail = True, ""
    if rank == 0:
        want = od.non_max_suppression(None, [t.to(dev) for t in whole], conf_thres=0.001, compat=False)
        if len(per) != B:
            ok, detail = False, f"{len(per)} images gathered, {B} expected"
        for i in range(B if ok else 0):
            a_, b_ = per[i], want[i]
            if (a_ is None) != (b_ is None) or (a_ is not None and not torch.equal(a_, b_)):
                ok, detail = False, f"image {i} differs"
                break
        ids = g.packed()[:, 7]
        if ok and not bool((ids[1:] >= ids[:-1]).all()):
            ok, detail = False, "global image ids not ascending"
        print(json.dumps({"ok": ok, "detail": detail, "world": world, "images": B, "rows": int(sum(g.totals)),
                          "kept_per_image": sum(g.totals) / B}))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()

"""Plain pinned host->device copy ceiling of this box for the headline batch (548 MB per rank), single process or one
process per GPU under torchrun — the ceiling `bench.py`'s e2e leg is judged against.

    python tools/h2d_bandwidth.py                                                        one GPU
    torchrun --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/h2d_bandwidth.py   N GPUs, all copying at once

Every rank copies its own pinned 548 MB buffer to its own GPU (8 chunks of 68.5 MB, like the 8-image chunks of the e2e
pipeline), 20 timed rounds after 3 warm-up rounds, all ranks starting each round from a barrier; the result is the
max-over-ranks time per round -> aggregate GB/s.  With --duplex the device->host copy of a 47 MB result (the padded
detection rows of a headline batch) runs on a second stream at the same time.  Prints one JSON line on rank 0."""
import argparse, json, os, statistics, sys
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--bytes", type=int, default=548352000)
ap.add_argument("--chunks", type=int, default=8)
ap.add_argument("--rounds", type=int, default=20)
ap.add_argument("--duplex", action="store_true")
ap.add_argument("--bind", action="store_true", help="bind the rank to the CPUs NVML reports as local to its GPU first")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if a.bind:
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    bench.bind_to_gpu_numa_node(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
n = a.bytes // 4
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device=dev)
views = list(zip(h.chunk(a.chunks), d.chunk(a.chunks)))
ho = torch.empty(47 * 1000 * 1000 // 4, dtype=torch.float32).pin_memory()
do = torch.empty_like(ho, device=dev)
s2 = torch.cuda.Stream(dev)


def round_():
    if a.duplex:
        with torch.cuda.stream(s2):
            ho.copy_(do, non_blocking=True)
    for x, y in views:
        y.copy_(x, non_blocking=True)
    if a.duplex:
        torch.cuda.current_stream(dev).wait_stream(s2)


ts = []
for i in range(3 + a.rounds):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    round_()
    e1.record()
    torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if i >= 3:
        ts.append(float(t.item()))
if rank == 0:
    ms = statistics.median(ts)
    print(json.dumps({"n_gpus": world, "bytes_per_rank": a.bytes, "chunks": a.chunks, "duplex": a.duplex, "bound_to_numa": a.bind,
                      "ms_per_round_max_over_ranks": ms, "per_rank_GBps": a.bytes / ms / 1e6, "aggregate_GBps": world * a.bytes / ms / 1e6,
                      "images_per_s_ceiling_headline": world * 64 / (ms * 1e-3), "host_cpus": os.cpu_count()}))
if world > 1:
    dist.destroy_process_group()

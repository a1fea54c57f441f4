"""Raw pinned host->device bandwidth of this box for the headline batch (548 MB), in 1 / 8 / 64 copies: the floor of bench.py e2e."""
import torch, time
dev = "cuda:0"
n = 548352000 // 4
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device=dev)
for chunks in (1, 8, 64):
    views = list(zip(h.chunk(chunks), d.chunk(chunks)))
    for _ in range(2):
        for a, b in views:
            b.copy_(a, non_blocking=True)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for a, b in views:
            b.copy_(a, non_blocking=True)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[2]
    print(f"H2D 548 MB in {chunks} copies: {ms:.2f} ms = {548.352 / ms:.1f} GB/s")

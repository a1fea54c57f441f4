"""Robustness fuzz: heads made of arbitrary bit patterns (NaN, inf, denormals, huge and negative sizes) must never hang
the pipeline.  Run under `timeout`; prints one line per case."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import objectdetectionpl_b200 as od

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
cases = [(2, 3, 20, [20, 10, 5]), (2, 3, 4, [40, 20, 10]), (1, 3, 80, [80, 40, 20]), (2, 3, 2, [24, 12, 6])]
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 12):
    B, A, C, grids = cases[it % len(cases)]
    mode = it % 3
    levels = []
    for G in grids:
        bits = torch.randint(-2**31, 2**31 - 1, (B, A * (5 + C), G, G), device=dev, generator=g, dtype=torch.int64).to(torch.int32)
        t = bits.view(torch.float32).clone()
        v = t.view(B, A, 5 + C, G, G)
        if mode == 1:      # plausible boxes, garbage conf / classes
            v[:, :, 0:2] = torch.rand(B, A, 2, G, G, device=dev, generator=g) * 100
            v[:, :, 2:4] = torch.rand(B, A, 2, G, G, device=dev, generator=g) * 50
        if mode == 2:      # garbage boxes, plausible conf / classes
            v[:, :, 4:] = torch.rand(B, A, 1 + C, G, G, device=dev, generator=g)
        levels.append(t.contiguous())
    t0 = time.time()
    rows, index, count = od.yolo_nms_raw(levels, A, want_index=True)
    torch.cuda.synchronize()
    print(f"case {it} mode {mode} B={B} C={C} grids={grids}: kept {count.tolist()} in {1e3 * (time.time() - t0):.1f} ms", flush=True)
print("fuzz ok")

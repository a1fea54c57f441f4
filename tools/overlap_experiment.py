"""Does running the pipeline on image sub-batches in several streams overlap the HBM-bound decode with the ALU-bound NMS?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import objectdetectionpl_b200 as od
from objectdetectionpl_b200 import synth

dev = torch.device("cuda:0")
B, C = 64, 80
g = torch.Generator(device=dev).manual_seed(1)
levels = []
for G in (80, 40, 20):
    t = torch.rand(B, 3, 5 + C, G, G, device=dev, generator=g)
    t[:, :, 0:2] *= 640; t[:, :, 2:4] = 8 + t[:, :, 2:4] * 152
    levels.append(t.view(B, 3 * (5 + C), G, G))

def run(nsplit, iters=30):
    streams = [torch.cuda.Stream() for _ in range(nsplit)]
    per = B // nsplit
    parts = [[t[i * per:(i + 1) * per] for t in levels] for i in range(nsplit)]
    def once():
        cur = torch.cuda.current_stream()
        for s, p in zip(streams, parts):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                od.yolo_nms_raw(p, 3)
        for s in streams:
            cur.wait_stream(s)
    for _ in range(5):
        once()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        once()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3

for n in (1, 2, 4, 8):
    us = run(n)
    print(f"{n} stream(s): {us:.1f} us per 64 images -> {64 / us * 1e6:.0f} img/s")

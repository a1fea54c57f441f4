import os, sys, time
sys.path.insert(0, "/root/repo")
import torch
import objectdetectionpl_b200 as od
from objectdetectionpl_b200 import synth
DEV = torch.device("cuda:0")
B, C = 64, 80
tg = synth.labels(B, C, 4).to(DEV)
stride = torch.tensor([8., 16., 32.])
anchors = (torch.tensor(synth.YOLOV5_ANCHORS).float().view(3, -1, 2) / stride.view(-1, 1, 1)).to(DEV)
p = [torch.randn(B, 3, 640 // s, 640 // s, 5 + C, device=DEV, requires_grad=True) for s in (8, 16, 32)]
def step():
    for t in p:
        t.grad = None
    od.v5_loss(p, tg, anchors, 3, 3, C)["loss"].backward()
for _ in range(20): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
N = 200
t0 = time.perf_counter(); e0.record()
for _ in range(N): step()
e1.record(); t1 = time.perf_counter()
torch.cuda.synchronize(); t2 = time.perf_counter()
print("host enqueue per step %.1f us; device per step %.1f us; wall incl. drain %.1f us" % ((t1 - t0) / N * 1e6, e0.elapsed_time(e1) / N * 1e3, (t2 - t0) / N * 1e6))
# forward only / backward only
import torch.autograd.profiler as prof
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA]) as pr:
    for _ in range(20): step()
    torch.cuda.synchronize()
print(pr.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))

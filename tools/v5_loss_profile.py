"""ncu target: one forward + backward of the fused v5 loss at BASELINE config 4 (B=64, C=80, 640x640)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import objectdetectionpl_b200 as od
from objectdetectionpl_b200 import synth
DEV = torch.device("cuda:0")
B, C = 64, 80
tg = synth.labels(B, C, 4).to(DEV)
stride = torch.tensor([8., 16., 32.])
anchors = (torch.tensor(synth.YOLOV5_ANCHORS).float().view(3, -1, 2) / stride.view(-1, 1, 1)).to(DEV)
p = [torch.randn(B, 3, 640 // s, 640 // s, 5 + C, device=DEV, requires_grad=True) for s in (8, 16, 32)]
for it in range(3):
    for t in p:
        t.grad = None
    od.v5_loss(p, tg, anchors, 3, 3, C)["loss"].backward()
torch.cuda.synchronize()
print("done")

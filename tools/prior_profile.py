import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import objectdetectionpl_b200 as od
from objectdetectionpl_b200 import synth
DEV = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "retina"
pri = synth.retina_priors(800) if which == "retina" else synth.ssd_priors()
loc, cls = synth.prior_heads(32, pri.shape[0], 80, 3)
loc, cls, pr = loc.to(DEV), cls.to(DEV), pri.to(DEV)
for _ in range(3):
    od.prior_nms_raw(loc, cls, pr)
torch.cuda.synchronize()

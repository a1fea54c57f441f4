"""GPU, N >= 2 devices on the box (skipped on a single-GPU box; run with `gpurun --gpus 2|8`): one process per GPU over
NCCL — the sharded run + detection all-gather equals the single-GPU run (tools/gather_check.py), and one process driving
two devices works (per-device kernel attributes)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NGPU = torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(NGPU < 2, reason="needs at least two GPUs on the box")
def test_nccl_gathered_set_equals_single_gpu_set_on_crowd_data():
    n = 8 if NGPU >= 8 else (4 if NGPU >= 4 else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "gather_check.py"), "--per-rank", "2", "--img", "640"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["ok"] and res["world"] == n and res["images"] == 2 * n and res["rows"] > 0, res


@pytest.mark.skipif(NGPU < 2, reason="needs at least two GPUs on the box")
def test_one_process_two_devices():
    """cluster_sort_kernel's > 48 KB dynamic shared memory opt-in is a per-device attribute: the first launch on a second
    device of the same process used to fail (cached in a process-wide static)."""
    import objectdetectionpl_b200 as od
    from objectdetectionpl_b200 import synth
    lv = synth.yolo_planar(2, 3, 5, [20, 10, 5], 160, 9, v5_view=True)
    outs = []
    for d in ("cuda:0", "cuda:1"):
        with torch.cuda.device(d):
            outs.append(od.non_max_suppression(None, [t.to(d) for t in lv]))
    host = od.non_max_suppression_host(None, [t.pin_memory() for t in lv], device="cuda:1")
    for a, b, h in zip(outs[0], outs[1], host):
        assert torch.equal(a.cpu(), b.cpu()) and torch.equal(a.cpu(), h)

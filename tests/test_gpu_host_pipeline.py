"""GPU: non_max_suppression_host (pinned host input, chunked copy/compute overlap) returns exactly what
non_max_suppression returns for the same heads on the device."""
import pytest
import torch

import objectdetectionpl_b200 as od
from objectdetectionpl_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("B,chunk,pin", [(5, 2, True), (8, 8, True), (3, 4, False), (9, 4, True)])
def test_host_pipeline_equals_device_api(B, chunk, pin):
    levels = synth.yolo_planar(B=B, A=3, C=20, grids=[20, 10, 5], img=160, seed=31 + B, v5_view=True)
    host = [t.pin_memory() if pin else t for t in levels]
    want, widx = od.non_max_suppression(None, [t.to(DEV) for t in levels], return_index=True)
    for rep in range(2):                       # second call reuses the cached buffers and streams
        got, gidx = od.non_max_suppression_host(None, host, device=DEV, chunk_images=chunk, return_index=True)
        assert len(got) == B
        for b in range(B):
            assert not got[b].is_cuda
            assert torch.equal(got[b], want[b].cpu()), f"image {b}"
            assert torch.equal(gidx[b], widx[b].cpu()), f"image {b}"


def test_host_pipeline_rejects_cuda_and_noncontiguous():
    levels = synth.yolo_planar(B=2, A=3, C=4, grids=[8], img=64, seed=1)
    with pytest.raises(TypeError):
        od.non_max_suppression_host(None, [t.to(DEV) for t in levels])
    with pytest.raises(TypeError):
        od.non_max_suppression_host(None, [levels[0].transpose(2, 3)])


def test_host_pipeline_sparse_none_entries():
    levels = synth.yolo_planar(B=4, A=3, C=4, grids=[8], img=64, seed=2)
    for t in levels:
        t.view(4, 3, 9, 8, 8)[1, :, 4] = 0.0          # image 1: nothing above the threshold
    got = od.non_max_suppression_host(None, levels, conf_thres=0.5, compat=False, device=DEV, chunk_images=3)
    want = od.non_max_suppression(None, [t.to(DEV) for t in levels], conf_thres=0.5, compat=False)
    assert got[1] is None and want[1] is None
    for b in (0, 2, 3):
        assert torch.equal(got[b], want[b].cpu())


def test_host_pipeline_async_two_batches_in_flight():
    """Batch i+1 submitted before batch i is collected: different inputs, same shapes (same cached buffers and streams);
    the rows of a call stay valid until the second submission after it."""
    batches = [[t.pin_memory() for t in synth.yolo_planar(B=6, A=3, C=20, grids=[20, 10, 5], img=160, seed=70 + i, v5_view=True)]
               for i in range(5)]
    wants = [od.non_max_suppression(None, [t.to(DEV) for t in lv], return_index=True) for lv in batches]
    handles, results = [], []
    for i, lv in enumerate(batches):
        handles.append(od.non_max_suppression_host_async(None, lv, device=DEV, chunk_images=4, return_index=True))
        if i >= 1:
            got, gidx = handles[i - 1].result()
            # compare right away: these rows live in a pinned buffer that the submission after next will overwrite
            for b in range(6):
                assert torch.equal(got[b], wants[i - 1][0][b].cpu()), f"batch {i - 1} image {b}"
                assert torch.equal(gidx[b], wants[i - 1][1][b].cpu())
            results.append(i - 1)
    got, gidx = handles[-1].result()
    assert handles[-1].result() is handles[-1].result()          # idempotent
    for b in range(6):
        assert torch.equal(got[b], wants[-1][0][b].cpu())
    assert results == [0, 1, 2, 3]

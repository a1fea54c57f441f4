"""CPU: the synthetic generators are deterministic, produce tie-free scores and well-formed layouts."""
import torch

from objectdetectionpl_b200 import synth
from oracle import ref_port as rp


def test_yolo_planar_deterministic_tie_free_and_terminating():
    a = synth.yolo_planar(2, 3, 20, [20, 10, 5], 160, seed=3, v5_view=True)
    b = synth.yolo_planar(2, 3, 20, [20, 10, 5], 160, seed=3, v5_view=True)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    assert [tuple(t.shape) for t in a] == [(2, 3, 20, 20, 25), (2, 3, 10, 10, 25), (2, 3, 5, 5, 25)]
    rows = rp.yolo_rows_from_planar(a, 3)
    score = rows[..., 4] * rows[..., 5:].max(-1)[0]
    for i in range(2):
        assert torch.unique(score[i]).numel() == score.shape[1]          # no duplicate fp32 scores
    assert float(rows[..., 2:4].min()) > -1.0                            # the reference loop terminates
    c = synth.yolo_planar(2, 3, 20, [20, 10, 5], 160, seed=4, v5_view=True)
    assert not torch.equal(a[0], c[0])


def test_views_share_planar_storage():
    v5 = synth.yolo_planar(1, 3, 4, [8], 64, seed=1, v5_view=True)[0]
    v3 = synth.yolo_planar(1, 3, 4, [8], 64, seed=1, v5_view=False)[0]
    assert tuple(v3.shape) == (1, 27, 8, 8)
    assert torch.equal(v5.reshape(-1), v3.reshape(-1))


def test_labels_and_crowd():
    t = synth.labels(8, 80, seed=4)
    assert t.shape[1] == 6 and t[:, 0].max() == 7 and bool((t[:, 0][1:] >= t[:, 0][:-1]).all())
    assert float(t[:, 2:4].min()) >= 0.05 and float(t[:, 4:6].max()) <= 0.31
    lv = synth.yolo_crowd(1, 3, 2, [16, 8], 128, seed=5, blobs=4, keep_frac=0.3)
    conf = lv[0].view(1, 3, 7, 16, 16)[:, :, 4]
    frac = float((conf >= 0.001).float().mean())
    assert 0.15 < frac < 0.45


def test_grids():
    assert synth.grids_for("yolov5", 640) == [80, 40, 20]
    assert synth.grids_for("yolov3", 416) == [13, 26, 52]
    assert synth.grids_for("yolov2", 416) == [13]
    assert sum(3 * g * g for g in synth.grids_for("yolov5", 640)) == 25200

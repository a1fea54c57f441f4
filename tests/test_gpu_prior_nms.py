"""GPU parity: SSD / RetinaNet prior decode + top-k class-agnostic NMS (model/SSD.py:249-310) against the
golden vectors of the unmodified reference and against the CPU oracle, quirks included."""
import pytest
import torch

import objectdetectionpl_b200 as od
from objectdetectionpl_b200 import synth
from objectdetectionpl_b200.postprocess import prior_nms_raw
from oracle import ref_port as rp
from tests.golden_io import load, T, unpack_list, SSD_CASES

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class _Self:
    def __init__(self, priors):
        self.iou_boxes = priors


def _close(got, want, what):
    assert len(got) == len(want)
    for i, (g, w) in enumerate(zip(got, want)):
        g = g.cpu()
        assert g.shape == w.shape, f"{what}[{i}]: {tuple(g.shape)} vs {tuple(w.shape)}"
        assert torch.equal(g[:, 4], w[:, 4]) and torch.equal(g[:, 6], w[:, 6]), f"{what}[{i}]: label column"
        torch.testing.assert_close(g[:, :4], w[:, :4], rtol=1e-5, atol=1e-5, msg=lambda m: f"{what}[{i}] boxes: {m}")
        torch.testing.assert_close(g[:, 5], w[:, 5], rtol=1e-5, atol=1e-6, msg=lambda m: f"{what}[{i}] scores: {m}")


@pytest.mark.parametrize("name", SSD_CASES)
def test_golden_reference_vectors(name):
    d = load(name)
    loc, cls, pri = T(d["loc"]).to(DEV), T(d["cls"]).to(DEV), T(d["priors"]).to(DEV)
    got = od.prior_non_max_suppression(_Self(pri), (loc, cls))
    _close(got, unpack_list(d, "out"), name)
    got = od.prior_non_max_suppression(_Self(pri), (loc, cls), topk=50, nms_thresh=0.3, class_thresh=0.3, mode="min")
    _close(got, unpack_list(d, "out_min"), name + "/min")


@pytest.mark.parametrize("P,C,topk,seed,mean", [(8732, 80, 100, 301, -3.0), (3069, 6, 700, 302, -1.0), (1000, 3, 100, 303, -2.0)])
def test_against_oracle(P, C, topk, seed, mean):
    pri = synth.ssd_priors() if P == 8732 else (synth.retina_priors(128) if P == 3069 else synth.ssd_priors()[:P])
    loc, cls = synth.prior_heads(2, P, C, seed, cls_mean=mean)
    want, widx = rp.ssd_nms(loc, cls, pri, topk=topk, return_index=True)
    got, gidx = od.prior_non_max_suppression(_Self(pri.to(DEV)), (loc.to(DEV), cls.to(DEV)), topk=topk, return_index=True)
    _close(got, want, f"seed{seed}")
    for a, b in zip(gidx, widx):
        assert torch.equal(a.cpu(), b)


def test_non_compat_mode_keeps_last_and_uses_own_boxes():
    pri = synth.ssd_priors()
    loc, cls = synth.prior_heads(1, pri.shape[0], 4, 310, cls_mean=-5.0)
    got, gidx = od.prior_non_max_suppression(_Self(pri.to(DEV)), (loc.to(DEV), cls.to(DEV)), compat=False, return_index=True)
    boxes = rp.prior_decode(loc[0], pri)
    score, label = cls[0].sigmoid().max(1)
    g, gi = got[0].cpu(), gidx[0].cpu()
    assert g.shape[0] >= 1
    torch.testing.assert_close(g[:, :4], boxes[gi], rtol=1e-5, atol=1e-6)
    assert torch.equal(g[:, 6], label[gi].float())
    torch.testing.assert_close(g[:, 5], score[gi], rtol=1e-5, atol=1e-6)
    assert bool((g[:-1, 5] >= g[1:, 5]).all())


def test_quirks_zero_and_one_candidate():
    pri = synth.ssd_priors()[:64].to(DEV)
    loc = torch.zeros(1, 64, 4, device=DEV)
    cls = torch.full((1, 64, 3), -9.0, device=DEV)
    out = od.prior_non_max_suppression(_Self(pri), (loc, cls))
    assert out[0].shape == (0, 7)
    cls[0, 5, 1] = 4.0
    with pytest.raises(IndexError):
        od.prior_non_max_suppression(_Self(pri), (loc, cls))
    out = od.prior_non_max_suppression(_Self(pri), (loc, cls), compat=False)
    assert out[0].shape == (1, 7) and out[0][0, 6].item() == 1.0
    with pytest.raises(TypeError):
        od.prior_non_max_suppression(_Self(pri), (loc, cls), mode="bogus")


# ---- radix-select top-k (topk.cu) against the full score sort it replaces ------------------------------------------------
@pytest.mark.parametrize("P,C,B,topk,quant,mean", [
    (8732, 21, 4, 100, None, -1.0),        # SSD300-like: thousands of candidates per image
    (30000, 8, 3, 100, None, 1.0),         # nearly every prior is a candidate
    (5000, 4, 3, 100, 0.5, 0.0),           # logits quantised to 0.5: massive score ties, the tie order decides the top-k
    (600, 3, 3, 200, None, -3.5),          # fewer candidates than topk
    (4000, 5, 2, 1, None, 0.0),            # topk = 1
    (9000, 6, 2, 1000, 0.25, 0.5),         # large topk with ties
])
def test_topk_select_equals_full_sort(P, C, B, topk, quant, mean, monkeypatch):
    g = torch.Generator().manual_seed(P + topk)
    pri = torch.rand(P, 4, generator=g) * 0.5 + 0.1
    loc = torch.randn(B, P, 4, generator=g) * 0.2
    cls = torch.randn(B, P, C, generator=g) * 2 + mean
    if quant:
        cls = torch.round(cls / quant) * quant
    args = (loc.to(DEV), cls.to(DEV), pri.to(DEV))
    got = prior_nms_raw(*args, topk=topk, want_index=True, compat=False)
    torch.cuda.synchronize()
    monkeypatch.setenv("B200DET_TOPK", "sort")
    want = prior_nms_raw(*args, topk=topk, want_index=True, compat=False)
    torch.cuda.synchronize()
    assert torch.equal(got[2], want[2])
    for b, k in enumerate(got[2][0].tolist()):
        assert torch.equal(got[1][b, :k], want[1][b, :k]), f"image {b}: kept priors differ"
        assert torch.equal(got[0][b, :k], want[0][b, :k]), f"image {b}: rows differ"

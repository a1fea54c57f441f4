"""Loader for tests/golden/*.npz (written by oracle/gen_golden.py from the unmodified reference)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def unpack_list(d, prefix):
    out = []
    for i in range(int(d[prefix + "_n"])):
        out.append(None if bool(d[f"{prefix}_{i}_none"]) else T(d[f"{prefix}_{i}"]))
    return out


def yolo_levels(d):
    return [T(d[f"level_{i}"]) for i in range(int(d["n_levels"]))]


YOLO_CASES = ["yolo_v5_tiny", "yolo_v5_mid", "yolo_v3_mid", "yolo_v2_g13", "yolo_v4_filter"]
SSD_CASES = ["ssd300_c5", "retina128_c6", "ssd300_sparse"]


def assert_boxes_close(got, want, rtol=1e-5, cap=None, what=""):
    """Box coordinates [K,4] within `rtol` RELATIVE fp32 tolerance (north-star: 1e-5).  The reference of a coordinate's
    magnitude is the largest |coordinate| of its row (floor 0.1): a merged x1 = sum(conf * x1) / sum(conf) over members
    with x1 in [-50, 50] can itself be ~0, and its rounding error — the two implementations only differ in the fp32
    summation order — is a few ulp of the TERMS, not of the result, so a bound relative to the coordinate alone is not
    attainable by any implementation (including the reference against itself under a different `sum` order).  There is
    no additive absolute term (round 1 used 1e-4 + 1e-5 |x|)."""
    got, want = got.detach().cpu().double(), want.detach().cpu().double()
    assert tuple(got.shape) == tuple(want.shape), f"{what}: shape {tuple(got.shape)} vs {tuple(want.shape)}"
    if want.numel() == 0:
        return
    scale = want.abs().amax(dim=-1, keepdim=True).clamp_min(0.1)
    tol = rtol * scale
    if cap is not None:
        tol = tol.clamp_max(cap)
    tol = tol.expand_as(want)
    err = (got - want).abs()
    bad = ~((err <= tol) | ((got == want) | (got.isnan() & want.isnan())))
    if bad.any():
        i = bad.nonzero()[0].tolist()
        raise AssertionError(f"{what}: box coordinate {i}: got {got[tuple(i)].item()!r} want {want[tuple(i)].item()!r} "
                             f"(err {err[tuple(i)].item():.3e} > tol {tol[tuple(i)].item():.3e}); {int(bad.sum())} bad of {bad.numel()}")


def assert_rows_close(got, want, rtol=1e-5, atol=1e-4, exact_cols=(4, 5, 6), what=""):
    """Detection rows [K,7]: cols 4..6 (conf, cls_conf, cls_id) bit-exact, box cols within 1e-5 relative (see
    assert_boxes_close; `atol` is kept in the signature for the callers and no longer used)."""
    assert (got is None) == (want is None), what
    if want is None:
        return
    assert tuple(got.shape) == tuple(want.shape), f"{what}: shape {tuple(got.shape)} vs {tuple(want.shape)}"
    got, want = got.detach().cpu(), want.detach().cpu()
    for c in exact_cols:
        bad = (got[:, c] != want[:, c]).nonzero().flatten()
        assert bad.numel() == 0, f"{what}: col {c} differs at rows {bad[:8].tolist()}"
    assert_boxes_close(got[:, :4], want[:, :4], rtol=rtol, what=what)

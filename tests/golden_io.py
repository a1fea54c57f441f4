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


def assert_rows_close(got, want, rtol=1e-5, atol=1e-4, exact_cols=(4, 5, 6), what=""):
    """Detection rows [K,7]: cols 4..6 (conf, cls_conf, cls_id) bit-exact, box cols within tolerance."""
    assert (got is None) == (want is None), what
    if want is None:
        return
    assert tuple(got.shape) == tuple(want.shape), f"{what}: shape {tuple(got.shape)} vs {tuple(want.shape)}"
    got, want = got.detach().cpu(), want.detach().cpu()
    for c in exact_cols:
        bad = (got[:, c] != want[:, c]).nonzero().flatten()
        assert bad.numel() == 0, f"{what}: col {c} differs at rows {bad[:8].tolist()}"
    torch.testing.assert_close(got[:, :4], want[:, :4], rtol=rtol, atol=atol, msg=lambda m: f"{what}: {m}")

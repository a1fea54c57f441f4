"""TEST INFRASTRUCTURE ONLY — ctypes wrapper of oracle/c/nms_oracle.c (full-size checker for the YOLO NMS)."""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libnms_oracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.run(["make", "-C", os.path.join(_HERE, "c")], check=True)
        _lib = ctypes.CDLL(_SO)
        _lib.yolo_nms_oracle_batch.restype = None
        _lib.yolo_nms_oracle_batch.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                               ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return _lib


def yolo_nms_rows(rows: torch.Tensor, conf_thres: float = -0.0151, nms_thres: float = 0.4):
    """rows [B,N,5+C] fp32 (xywh) -> (list of [K,7] or None, list of kept candidate indices or None)."""
    lib = load()
    rows = rows.contiguous().float()
    B, N, F = rows.shape
    out = np.empty((B, N, 7), np.float32)
    idx = np.empty((B, N), np.int32)
    cnt = np.empty((B,), np.int32)
    lib.yolo_nms_oracle_batch(rows.data_ptr(), B, N, F, conf_thres, nms_thres, out.ctypes.data, idx.ctypes.data, cnt.ctypes.data)
    dets, inds = [], []
    for b in range(B):
        k = int(cnt[b])
        dets.append(torch.from_numpy(out[b, :k].copy()) if k else None)
        inds.append(torch.from_numpy(idx[b, :k].astype(np.int64)) if k else None)
    return dets, inds

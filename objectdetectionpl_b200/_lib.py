"""ctypes binding of libb200det.so (C ABI in include/b200det.h).  No CPU fallback: if the shared library
is missing or there is no CUDA device the product path raises — it never routes through the oracle."""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_size_t, c_uint8, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200det.so")

MAX_LEVELS, MAX_ANCHORS, MAX_CLASSES, MAX_CANDIDATES, TILE = 8, 16, 4095, 1 << 20, 512
DECODE_NONE, DECODE_YOLO_EXP, DECODE_YOLOV5, DECODE_YOLOV4_NORM = 0, 1, 2, 3
LAYOUT_PLANAR, LAYOUT_CHANNELS_LAST = 0, 1
IOU, GIOU, DIOU, CIOU = 0, 1, 2, 3


class YoloDesc(Structure):
    _fields_ = [("batch", c_int32), ("num_anchors", c_int32), ("num_classes", c_int32), ("num_levels", c_int32),
                ("head", c_void_p * MAX_LEVELS), ("grid", c_int32 * MAX_LEVELS), ("decode_mode", c_int32),
                ("stride", c_float * MAX_LEVELS), ("anchors", ((c_float * 2) * MAX_ANCHORS) * MAX_LEVELS),
                ("conf_thres", c_float), ("nms_thres", c_float), ("layout", c_int32), ("scale_x_y", c_float)]


class PriorDesc(Structure):
    _fields_ = [("batch", c_int32), ("num_priors", c_int32), ("num_classes", c_int32), ("loc", c_void_p),
                ("cls", c_void_p), ("priors", c_void_p), ("topk", c_int32), ("nms_thresh", c_float),
                ("class_thresh", c_float), ("mode_min", c_int32), ("compat", c_int32)]


class V5Level(Structure):
    _fields_ = [("pi", c_void_p), ("batch", c_int32), ("na", c_int32), ("ny", c_int32), ("nx", c_int32), ("fields", c_int32),
                ("b", c_void_p), ("a", c_void_p), ("gj", c_void_p), ("gi", c_void_p), ("tcls", c_void_p), ("tbox", c_void_p),
                ("anch", c_void_p), ("m_dev", c_void_p), ("gpi", c_void_p)]


_vp, _i32, _i64, _f, _sz = c_void_p, c_int32, c_int64, c_float, c_size_t
_PY, _PP = POINTER(YoloDesc), POINTER(PriorDesc)

# name -> (restype, argtypes).  Kept in sync with include/b200det.h (tests/test_cabi.py checks that every
# symbol declared there is listed here and exported by the library).
SIGNATURES = {
    "b200det_version": (_i32, []),
    "b200det_last_error": (c_char_p, []),
    "b200det_yolo_num_candidates": (_i32, [_PY, POINTER(_i32), POINTER(_i32)]),
    "b200det_yolo_workspace_bytes": (_sz, [_PY]),
    "b200det_yolo_nms": (_i32, [_PY, _vp, _sz, _vp, _vp, _vp, _vp]),
    "b200det_yolo_nms_early": (_i32, [_PY, _vp, _sz, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200det_yolo_nms_packed": (_i32, [_PY, _vp, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200det_yolo_stage_emit_packed": (_i32, [_PY, _vp, _sz, _vp, _vp, _vp, _vp, _vp]),
    "b200det_yolo_stage_reset": (_i32, [_PY, _vp, _sz, _vp]),
    "b200det_yolo_stage_decode": (_i32, [_PY, _vp, _sz, _vp]),
    "b200det_yolo_stage_sort": (_i32, [_PY, _vp, _sz, _vp]),
    "b200det_yolo_stage_nms": (_i32, [_PY, _vp, _sz, _vp]),
    "b200det_yolo_stage_emit": (_i32, [_PY, _vp, _sz, _vp, _vp, _vp, _vp]),
    "b200det_yolo_workspace_field": (_i32, [_PY, c_char_p, POINTER(_sz), POINTER(_sz)]),
    "b200det_decode_box": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _f, _vp, _vp]),
    "b200det_yolo_forward_dynamic": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _f, _vp, _i64, _vp, _i64, _vp, _i64, _vp]),
    "b200det_prior_workspace_bytes": (_sz, [_PP]),
    "b200det_prior_nms": (_i32, [_PP, _vp, _sz, _vp, _vp, _vp, _vp, _vp]),
    "b200det_prior_stage_decode": (_i32, [_PP, _vp, _sz, _vp]),
    "b200det_prior_stage_select_nms": (_i32, [_PP, _vp, _sz, _vp, _vp, _vp, _vp, _vp]),
    "b200det_xywh2xyxy": (_i32, [_vp, _vp, _i64, _vp]),
    "b200det_bbox_iou_plus1": (_i32, [_vp, _i64, _vp, _i64, _i32, _vp, _vp]),
    "b200det_pair_iou": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "b200det_bbox_iou_v5_fwd": (_i32, [_vp, _i64, _i64, _vp, _i64, _i64, _i64, _i32, _i32, _vp, _vp]),
    "b200det_bbox_iou_v5_bwd": (_i32, [_vp, _i64, _i64, _vp, _i64, _i64, _i64, _i32, _i32, _vp, _vp, _vp]),
    "b200det_build_targets_v5_level": (_i32, [_vp, _i32, POINTER(_f), _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp,
                                              _vp, _vp, _vp]),
    "b200det_build_targets_v5": (_i32, [_vp, _i32, _i32, POINTER(_f), _i32, POINTER(_i32), POINTER(_i32)] +
                                 [POINTER(_vp)] * 7 + [_vp, _vp]),
    "b200det_v5_match_fwd": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp]),
    "b200det_v5_match_bwd": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp]),
    "b200det_v5_loss_fwd": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _f, _f, _f, _f,
                                   _i32, _vp, _vp, _vp, _vp]),
    "b200det_v5_loss_combine": (_i32, [_vp, _i32, _f, _f, _f, _vp, _vp]),
    "b200det_v5_loss_combine_bwd": (_i32, [_vp, _vp, _vp, _vp, _f, _f, _f, _vp, _vp]),
    "b200det_v5_loss_bwd": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _f, _f, _f, _f,
                                   _i32, _vp, _vp, _f, _f, _f, _vp, _vp]),
    "b200det_v5_loss_bwd_full": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _f, _f, _f, _f,
                                        _i32, _vp, _vp, _f, _f, _f, _vp, _vp]),
    "b200det_v5_loss_fwd_all": (_i32, [_vp, _i32, _i32, _f, _f, _f, _f, _i32, _f, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200det_v5_loss_bwd_all": (_i32, [_vp, _i32, _i32, _f, _f, _f, _f, _i32, _f, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200det_build_targets_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "b200det_build_targets": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _f, _vp, _sz] + [_vp] * 10 + [_vp]),
    "b200det_ssd_match_workspace_bytes": (_sz, [_i32, _i32]),
    "b200det_ssd_match": (_i32, [_vp, _i32, _vp, _i32, _f, _vp, _sz, _vp, _vp, _vp]),
    "b200det_retina_assign_workspace_bytes": (_sz, [_i32, _i32]),
    "b200det_retina_assign": (_i32, [_vp, _i32, _vp, _i32, _i32, _f, _vp, _sz, _vp, _vp, _vp]),
    "b200det_pack_detections": (_i32, [_vp, _vp, _i32, _i64, _i32, _i32, _vp, _i64, _vp, _vp]),
    "b200det_batch_statistics_workspace_bytes": (_sz, [_i32, _i32]),
    "b200det_batch_statistics": (_i32, [_vp, _vp, _vp, _i32, _i32, _vp, _i32, _f, _vp, _sz, _vp, _vp]),
    "b200det_yolo_statistics_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "b200det_yolo_statistics_level": (_i32, [_vp, _i32, _i32, _i32, _i32, _vp, _f, _vp, _i32, _f, _vp, _sz, _vp, _vp, _vp]),
    "b200det_ap_per_class_workspace_bytes": (_sz, [_i32]),
    "b200det_ap_per_class": (_i32, [_vp, _vp, _vp, _i32, _vp, _vp, _i32, _vp, _sz, _vp, _vp, _vp, _vp, _vp]),
}

_lib = None
_lock = threading.Lock()


def load():
    """Load libb200det.so (built in-tree by `__graft_entry__.build()` / csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                                   "g.build()'` (or `make -C objectdetectionpl_b200/csrc`). There is no CPU fallback.")
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = res, args
            _lib = lib
    return _lib


_glue = None


def hostglue():
    """The host-glue module (csrc/hostglue.cpp, built next to libb200det.so): the tail of the Python wrappers that runs
    between the counts event and the return.  Host-only; like the library itself it has no stand-in."""
    global _glue
    if _glue is None:
        try:
            from . import _hostglue
        except ImportError as e:
            raise RuntimeError("objectdetectionpl_b200/_hostglue*.so not found or not loadable: build it with "
                               "`python -c 'import __graft_entry__ as g; g.build()'` "
                               f"(or `make -C objectdetectionpl_b200/csrc`) — {e}") from e
        _glue = _hostglue
    return _glue


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().b200det_last_error().decode("utf-8", "replace")
        if rc < 0:
            raise ValueError(f"b200det {what}: {msg} (code {rc})")
        raise RuntimeError(f"b200det {what}: {msg}")


def require_cuda(t: torch.Tensor, name: str, dtype=torch.float32):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{name} is on {t.device}: b200det has no CPU path (move the tensor to a CUDA device)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    return t


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


_ws_cache = {}


def workspace(nbytes: int, device, stream: int = None) -> torch.Tensor:
    """Stream-keyed scratch buffer (uint8), grown on demand and reused between calls on the same stream
    (`stream`: the current stream's handle if the caller already has it)."""
    dev = device if isinstance(device, torch.device) else torch.device(device)
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), stream_ptr(dev) if stream is None else stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)
        _ws_cache[key] = buf
    return buf

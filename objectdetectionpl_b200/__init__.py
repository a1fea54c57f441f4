"""objectdetectionpl_b200 — B200-native (sm_100a CUDA behind a C ABI) replacement for the detection
post-processing and target-assignment hot path of Leyan529/ObjectDetectionPL.

Public names mirror the reference (LightningFunc/accuracy.py, model/*.py::non_max_suppression):
    non_max_suppression, non_max_suppression_host (pinned host input, copy/compute overlap), non_max_suppression_v2,
    prior_non_max_suppression, decode_box,
    xywh2xyxy, bbox_iou, iou, bbox_iou_v5, build_targets, build_targets_v5, v5_match_level,
    ssd_match, retina_assign, get_batch_statistics, ap_per_class (test-time metrics), install (drop-in monkey patch),
    dist (image-sharded multi-GPU glue).
The CUDA library is loaded lazily on first use; importing the package needs neither a GPU nor the .so.
"""
from .boxes import bbox_iou, bbox_iou_v5, iou, xywh2xyxy
from .postprocess import (decode_box, get_region_boxes, yolo_forward_dynamic, clear_host_pipes, non_max_suppression, non_max_suppression_host, non_max_suppression_host_async,
                          non_max_suppression_v2,
                          prior_non_max_suppression, prior_non_max_suppression_host,
                          prior_nms_raw, yolo_nms_raw, YOLO_FORCED_CONF_THRES)
from .targets import (build_targets, build_targets_v5, retina_assign, ssd_match, v5_loss, v5_loss_level,
                      v5_match_level)
from .metrics import (ap_per_class, ap_per_class_device, batch_statistics_raw, get_batch_statistics, get_yolo_statistics,
                      yolo_statistics_level)
from .patch import install, install_losses, install_metrics, install_model
from . import dist, synth

__all__ = ["prior_non_max_suppression_host", "yolo_forward_dynamic", "get_region_boxes", "clear_host_pipes", "non_max_suppression", "non_max_suppression_host", "non_max_suppression_host_async", "non_max_suppression_v2", "prior_non_max_suppression", "decode_box", "xywh2xyxy",
           "bbox_iou", "iou", "bbox_iou_v5", "build_targets", "build_targets_v5", "v5_match_level", "v5_loss_level", "v5_loss", "ssd_match",
           "retina_assign", "get_batch_statistics", "get_yolo_statistics", "yolo_statistics_level", "ap_per_class", "batch_statistics_raw", "ap_per_class_device", "install",
           "install_losses", "install_metrics", "install_model", "dist", "synth", "yolo_nms_raw", "prior_nms_raw",
           "YOLO_FORCED_CONF_THRES"]
__version__ = "0.1.0"

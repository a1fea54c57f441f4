"""Drop-in installation: the same `setattr` monkey-patch idiom the reference uses for its own step
functions (model/YOLOV5.py:134-150) and the module-global lookups of LightningFunc/losses.py:5-6.
`LightningFunc/step.py` stays untouched."""
from __future__ import annotations

import sys

from . import boxes, metrics, postprocess, targets


def install_model(model_cls):
    """Replace `non_max_suppression` on a reference model class (YOLOv2/3/4/5, SSD, RetinaNet)."""
    name = model_cls.__name__.lower()
    if name in ("ssd", "retinanet"):
        fn = postprocess.prior_non_max_suppression
    elif name == "yolov2":
        fn = postprocess.non_max_suppression_v2
    else:
        fn = postprocess.non_max_suppression
    setattr(model_cls, "non_max_suppression", fn)
    if name in ("yolov2", "yolov3", "yolov4"):
        setattr(model_cls, "get_yolo_statistics", metrics.get_yolo_statistics)     # step.py:99
        # the model's constructor re-binds the attribute from ITS module's global (`__build_func`, model/YOLOV3.py:252:
        # `setattr(obj, "get_yolo_statistics", get_yolo_statistics)`), so that global is patched too — otherwise building
        # the model after install() would silently restore the stock function
        mod = sys.modules.get(getattr(model_cls, "__module__", None))
        if mod is not None and hasattr(mod, "get_yolo_statistics"):
            mod.get_yolo_statistics = metrics.get_yolo_statistics
    return model_cls


def install_losses(losses_module=None, accuracy_module=None):
    """Patch the target-assignment globals the reference's loss classes resolve at call time
    (`build_targets_v5`, `bbox_iou_v5`: losses.py:102,118) and the function the RegionLoss classes copy into
    `self.build_targets` (losses.py:492,654,815 — patch before constructing the criterion)."""
    if losses_module is not None:
        losses_module.build_targets_v5 = targets.build_targets_v5
        losses_module.bbox_iou_v5 = boxes.bbox_iou_v5
        losses_module.build_targets = targets.build_targets
        losses_module.bbox_iou = boxes.bbox_iou
        losses_module.iou = boxes.iou
    if accuracy_module is not None:
        accuracy_module.build_targets_v5 = targets.build_targets_v5
        accuracy_module.bbox_iou_v5 = boxes.bbox_iou_v5
        accuracy_module.build_targets = targets.build_targets
        accuracy_module.bbox_iou = boxes.bbox_iou
        accuracy_module.xywh2xyxy = boxes.xywh2xyxy
        accuracy_module.iou = boxes.iou


def install_metrics(step_module=None, accuracy_module=None):
    """Patch the test-time metrics: `LightningFunc/step.py:11` binds `get_batch_statistics` and `ap_per_class` into its own
    namespace at import, so both modules are patched."""
    for m in (step_module, accuracy_module):
        if m is not None:
            m.get_batch_statistics = metrics.get_batch_statistics
            m.ap_per_class = metrics.ap_per_class


def install(*model_classes, losses_module=None, accuracy_module=None, step_module=None):
    for c in model_classes:
        install_model(c)
    install_losses(losses_module, accuracy_module)
    install_metrics(step_module, accuracy_module)

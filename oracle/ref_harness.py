"""TEST INFRASTRUCTURE ONLY — loader for the *unmodified* reference, used in this container to pin the oracle.

`/root/reference` (Leyan529/ObjectDetectionPL) is pure Python, so it cannot be compiled into
`oracle/_ref`; instead it is imported here, in the build container, to
  (1) validate the restatement in `oracle/ref_port.py`, and
  (2) generate the golden vectors committed under `tests/golden/` (see `oracle/gen_golden.py`).
It does NOT travel to the GPU box (`/root/reference` does not exist there): nothing under `-m gpu`,
`smoke()` or `bench.py` imports this module.

The shims below only make the reference importable/runnable on a CPU-only, Lightning-less box
(SURVEY.md §8c); no reference source is copied or modified:
  * `pytorch_lightning`, `torchinfo`, `matplotlib(.pyplot)` are absent  -> stub modules;
  * `model/YOLOV3.py:3` uses `collections.Iterable` (removed in py3.10) -> alias;
  * class bodies `open("dataset//pallete")` relative to the repo root    -> chdir while importing;
  * `.cuda()` / `torch.cuda.FloatTensor` are hard-coded in SSD.py:305, accuracy.py:421 -> CPU aliases.
"""
import collections
import collections.abc
import contextlib
import importlib
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("B200DET_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "LightningFunc"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


_installed = False


def install_shims():
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    pl = _stub("pytorch_lightning", LightningModule=nn.Module)
    pl.LightningModule = nn.Module
    _stub("torchinfo", summary=lambda *a, **k: None)
    mpl = _stub("matplotlib")
    plt = _stub("matplotlib.pyplot")
    mpl.pyplot = plt
    if not hasattr(collections, "Iterable"):
        collections.Iterable = collections.abc.Iterable
    # CPU aliases for the .cuda()-hard-coded bits of the reference.
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.cuda.FloatTensor = torch.FloatTensor
        torch.cuda.LongTensor = torch.LongTensor
        torch.cuda.ByteTensor = torch.ByteTensor
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _installed = True


@contextlib.contextmanager
def _in_ref_root():
    cwd = os.getcwd()
    os.chdir(REF_ROOT)
    try:
        yield
    finally:
        os.chdir(cwd)


def ref_import(modname: str):
    """Import a module of the reference (e.g. 'LightningFunc.accuracy', 'model.YOLOV5')."""
    install_shims()
    with _in_ref_root():
        return importlib.import_module(modname)


# Convenience accessors -----------------------------------------------------------------------

def accuracy():
    return ref_import("LightningFunc.accuracy")


def losses():
    return ref_import("LightningFunc.losses")


def yolo_nms(version: int = 5):
    """Unbound reference `non_max_suppression` of YOLOv{2,3,4,5} (model/YOLOV*.py)."""
    mod = ref_import({2: "model.YOLOV2", 3: "model.YOLOV3", 4: "model.YOLOV4", 5: "model.YOLOV5"}[version])
    cls = getattr(mod, {2: "YOLOv2", 3: "YOLOv3", 4: "YOLOv4", 5: "YOLOv5"}[version])
    return cls.non_max_suppression


def ssd_nms(which: str = "SSD"):
    mod = ref_import("model." + which)
    return getattr(mod, which).non_max_suppression

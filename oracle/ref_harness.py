"""TEST INFRASTRUCTURE ONLY — loader for the *unmodified* reference (Leyan529/ObjectDetectionPL, pure Python).

In the build container it is imported from `/root/reference` to
  (1) validate the restatement in `oracle/ref_port.py`, and
  (2) generate the golden vectors committed under `tests/golden/` (see `oracle/gen_golden.py`).
On the GPU box `/root/reference` does not exist; there the byte-identical copy staged by `oracle/stage_ref.py` into the
git-ignored `oracle/_ref/` is imported instead (checked against `oracle/ref_manifest.json`), so that
  (3) the `-m gpu` drop-in tests can run the reference's own criterion / NMS / test_step stock and with
      `objectdetectionpl_b200.install()` applied, and
  (4) `bench.py --impl reference` and the `cpu_baseline` leg time the reference's own CPU code.
Nothing under `objectdetectionpl_b200/` imports this module: the product path never touches the reference.

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

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _find_root():
    env = os.environ.get("B200DET_REFERENCE_ROOT")
    for cand in ([env] if env else []) + ["/root/reference", _STAGED]:
        if cand and os.path.isdir(os.path.join(cand, "LightningFunc")):
            return cand
    return "/root/reference"


REF_ROOT = _find_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "LightningFunc"))


def is_staged_copy() -> bool:
    return os.path.abspath(REF_ROOT) == os.path.abspath(_STAGED)


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
    # CPU aliases for the .cuda()-hard-coded bits of the reference (on a GPU box: only inside `cpu_only()`).
    if not torch.cuda.is_available():
        _alias_cuda_to_cpu()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _installed = True


def _alias_cuda_to_cpu():
    saved = (torch.Tensor.cuda, torch.cuda.FloatTensor, torch.cuda.LongTensor, torch.cuda.ByteTensor)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.FloatTensor = torch.FloatTensor
    torch.cuda.LongTensor = torch.LongTensor
    torch.cuda.ByteTensor = torch.ByteTensor
    return saved


@contextlib.contextmanager
def cpu_only():
    """Run the reference's CPU path on a box that HAS a GPU: while active, `.cuda()` and the `torch.cuda.*Tensor`
    constructors the reference hard-codes (SSD.py:305, accuracy.py:421, losses.py:73-99) resolve to their CPU forms."""
    saved = _alias_cuda_to_cpu()
    try:
        yield
    finally:
        torch.Tensor.cuda, torch.cuda.FloatTensor, torch.cuda.LongTensor, torch.cuda.ByteTensor = saved


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

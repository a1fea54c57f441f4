"""CPU: the C-ABI library loads, exports every symbol include/b200det.h declares, the ctypes mirror of the
descriptor structs matches the C layout, and argument validation works without launching anything."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

from objectdetectionpl_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200det.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200det_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(L.LIB_PATH):
        sys.path.insert(0, ROOT)
        import __graft_entry__ as g
        g.build()
    return L.load()


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert n in L.SIGNATURES, f"{n} declared in b200det.h but missing from _lib.SIGNATURES"
        assert getattr(lib, n) is not None
    for n in L.SIGNATURES:
        assert n in names, f"{n} bound in _lib.py but not declared in b200det.h"


def test_version_and_limits_match_header(lib):
    src = open(HEADER).read()
    macro = lambda m: int(re.search(rf"#define {m}\s+\(?([0-9<< ]+)\)?", src).group(1).replace(" ", "").replace("1<<20", str(1 << 20)))
    assert lib.b200det_version() == macro("B200DET_VERSION")
    assert L.MAX_LEVELS == macro("B200DET_MAX_LEVELS") and L.MAX_ANCHORS == macro("B200DET_MAX_ANCHORS")
    assert L.MAX_CLASSES == macro("B200DET_MAX_CLASSES") and L.TILE == macro("B200DET_TILE")
    assert L.MAX_CANDIDATES == 1 << 20


def test_struct_layout_matches_c(tmp_path):
    prog = tmp_path / "layout.c"
    prog.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "b200det.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(b200det_yolo_desc), offsetof(b200det_yolo_desc, head),
         offsetof(b200det_yolo_desc, grid), offsetof(b200det_yolo_desc, decode_mode), offsetof(b200det_yolo_desc, stride),
         offsetof(b200det_yolo_desc, anchors), offsetof(b200det_yolo_desc, conf_thres), offsetof(b200det_yolo_desc, nms_thres),
         offsetof(b200det_yolo_desc, layout), offsetof(b200det_yolo_desc, scale_x_y));
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(b200det_prior_desc), offsetof(b200det_prior_desc, loc),
         offsetof(b200det_prior_desc, priors), offsetof(b200det_prior_desc, topk), offsetof(b200det_prior_desc, mode_min),
         offsetof(b200det_prior_desc, compat));
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(b200det_v5_level), offsetof(b200det_v5_level, batch),
         offsetof(b200det_v5_level, fields), offsetof(b200det_v5_level, b), offsetof(b200det_v5_level, tbox),
         offsetof(b200det_v5_level, m_dev), offsetof(b200det_v5_level, gpi));
  return 0;
}''')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    y = [int(v) for v in out[0].split()]
    Y = L.YoloDesc
    assert y == [ctypes.sizeof(Y), Y.head.offset, Y.grid.offset, Y.decode_mode.offset, Y.stride.offset, Y.anchors.offset,
                 Y.conf_thres.offset, Y.nms_thres.offset, Y.layout.offset, Y.scale_x_y.offset]
    # the binding INTEGRATION.md shows a maintainer: execute its class statement and compare it with the C layout
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"^class YoloDesc\(ctypes\.Structure\):.*?\n((?:[ \t]+.*\n)+)", doc, flags=re.M)
    assert m, "INTEGRATION.md no longer shows the YoloDesc binding"
    ns = {"ctypes": ctypes}
    exec(m.group(0), ns)
    D = ns["YoloDesc"]
    assert [f[0] for f in D._fields_] == [f[0] for f in Y._fields_], "INTEGRATION.md's YoloDesc fields differ from _lib.YoloDesc"
    assert ctypes.sizeof(D) == y[0]
    for name, _ in Y._fields_:
        assert getattr(D, name).offset == getattr(Y, name).offset and getattr(D, name).size == getattr(Y, name).size, name
    p = [int(v) for v in out[1].split()]
    P = L.PriorDesc
    assert p == [ctypes.sizeof(P), P.loc.offset, P.priors.offset, P.topk.offset, P.mode_min.offset, P.compat.offset]
    v = [int(x) for x in out[2].split()]
    V = L.V5Level
    assert v == [ctypes.sizeof(V), V.batch.offset, V.fields.offset, V.b.offset, V.tbox.offset, V.m_dev.offset, V.gpi.offset]


def _desc(B=64, A=3, C=80, grids=(80, 40, 20)):
    d = L.YoloDesc()
    d.batch, d.num_anchors, d.num_classes, d.num_levels = B, A, C, len(grids)
    for i, g in enumerate(grids):
        d.grid[i] = g
        d.head[i] = 0x1000          # never dereferenced by the query functions
    d.conf_thres, d.nms_thres = -0.0151, 0.4
    return d


def test_size_queries_and_validation_without_a_gpu(lib):
    d = _desc()
    n, n_pad = ctypes.c_int32(), ctypes.c_int32()
    assert lib.b200det_yolo_num_candidates(ctypes.byref(d), ctypes.byref(n), ctypes.byref(n_pad)) == 0
    assert (n.value, n_pad.value) == (25200, 26112)          # 19200 + 4800 -> 5120 + 1200 -> 1536: levels start on tile boundaries
    nbytes = lib.b200det_yolo_workspace_bytes(ctypes.byref(d))
    assert 100e6 < nbytes < 400e6
    # null workspace -> EINVAL, message available, nothing launched
    rc = lib.b200det_yolo_nms(ctypes.byref(d), None, 0, None, None, None, None)
    assert rc == -1 and b"workspace" in lib.b200det_last_error()
    # too many classes -> ELIMIT
    bad = _desc(C=5000)
    assert lib.b200det_yolo_num_candidates(ctypes.byref(bad), ctypes.byref(n), ctypes.byref(n_pad)) == 0
    rc = lib.b200det_yolo_nms(ctypes.byref(bad), ctypes.c_void_p(256), 1 << 40, None, None, None, None)
    assert rc == -2 and b"num_classes" in lib.b200det_last_error()
    # too many candidates per image -> ELIMIT
    huge = _desc(grids=(1024,))
    assert lib.b200det_yolo_num_candidates(ctypes.byref(huge), ctypes.byref(n), ctypes.byref(n_pad)) == -2
    # misaligned pointers are rejected before any launch
    assert lib.b200det_xywh2xyxy(ctypes.c_void_p(4), ctypes.c_void_p(16), 8, None) == -1
    assert lib.b200det_bbox_iou_plus1(ctypes.c_void_p(16), 3, ctypes.c_void_p(32), 8, 1, ctypes.c_void_p(64), None) == -1
    assert lib.b200det_bbox_iou_v5_fwd(ctypes.c_void_p(16), 8, 1, ctypes.c_void_p(32), 8, 1, 8, 0, 7, ctypes.c_void_p(64), None) == -1
    # workspace introspection
    off, nb = ctypes.c_size_t(), ctypes.c_size_t()
    assert lib.b200det_yolo_workspace_field(ctypes.byref(d), b"box4", ctypes.byref(off), ctypes.byref(nb)) == 0
    assert nb.value == 64 * 26112 * 16 and off.value % 256 == 0
    assert lib.b200det_yolo_workspace_field(ctypes.byref(d), b"nope", ctypes.byref(off), ctypes.byref(nb)) == -1
    assert lib.b200det_build_targets_workspace_bytes(64, 3, 52, 6400) >= 64 * 3 * 52 * 52 * 4
    assert lib.b200det_ssd_match_workspace_bytes(8732, 10) > 0 and lib.b200det_retina_assign_workspace_bytes(32, 100) > 0
    pd = L.PriorDesc()
    pd.batch, pd.num_priors, pd.num_classes, pd.topk = 32, 120087, 80, 100
    assert lib.b200det_prior_workspace_bytes(ctypes.byref(pd)) > 32 * 120087 * 16


def test_python_layer_refuses_cpu_tensors_and_bad_shapes():
    import torch
    import objectdetectionpl_b200 as od
    with pytest.raises(RuntimeError, match="no CPU path"):
        od.xywh2xyxy(torch.zeros(4, 4))
    with pytest.raises(RuntimeError, match="no CPU path"):
        od.build_targets_v5([(1, 3, 8, 8, 6)] * 3, torch.zeros(2, 6), torch.ones(3, 3, 2), 3, 3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        od.prior_non_max_suppression(type("S", (), {"iou_boxes": torch.zeros(8, 4)})(), (torch.zeros(1, 8, 4), torch.zeros(1, 8, 3)))
    with pytest.raises(TypeError):
        od.prior_non_max_suppression(None, (torch.zeros(1, 8, 4), torch.zeros(1, 8, 3)), mode="bogus")
    with pytest.raises(ValueError):
        od.non_max_suppression(None, [])


def test_slot_count_formula_of_the_host_pipeline_matches_the_library(lib):
    """postprocess.non_max_suppression_host sizes its pinned output with the per-level tile padding; the library is the
    authority (every level starts on a 512-slot tile boundary)."""
    for A, C, grids in ((3, 80, (80, 40, 20)), (3, 80, (13, 26, 52)), (5, 20, (13,)), (3, 4, (7, 5, 3)), (3, 1, (160, 80, 40))):
        d = L.YoloDesc()
        d.batch, d.num_anchors, d.num_classes, d.num_levels = 2, A, C, len(grids)
        for i, g in enumerate(grids):
            d.grid[i] = g
        n, n_pad = ctypes.c_int32(), ctypes.c_int32()
        assert lib.b200det_yolo_num_candidates(ctypes.byref(d), ctypes.byref(n), ctypes.byref(n_pad)) == 0
        assert n.value == sum(A * g * g for g in grids)
        assert n_pad.value == sum((A * g * g + L.TILE - 1) // L.TILE * L.TILE for g in grids)


def test_makefile_builds_every_source_the_library_and_the_host_glue_from_scratch():
    """`make -n -B` (dry run, everything out of date): one sm_100a nvcc line per .cu under csrc/, the -shared link of
    libb200det.so and the host-glue build — a fresh checkout must be buildable by `__graft_entry__.build()`."""
    import glob
    import subprocess
    csrc = os.path.join(ROOT, "objectdetectionpl_b200", "csrc")
    out = subprocess.run(["make", "-C", csrc, "-n", "-B"], capture_output=True, text=True, check=True).stdout
    lines = out.splitlines()
    for cu in sorted(glob.glob(os.path.join(csrc, "*.cu"))):
        name = os.path.basename(cu)
        assert any(f"-c {name} " in ln and "arch=compute_100a,code=sm_100a" in ln and "-lineinfo" in ln for ln in lines), name
    assert any("-shared" in ln and "libb200det.so" in ln for ln in lines)
    assert any("build_hostglue.py" in ln and "_hostglue" in ln for ln in lines)

"""TEST INFRASTRUCTURE ONLY — stages the UNMODIFIED reference for the GPU box.

`/root/reference` exists only in the build container.  The reference is ~4 k lines of MIT-licensed Python, so it cannot
be compiled into a binary, but it can travel: this recipe copies the files of the hot path and of the modules they
import, byte for byte, into the git-ignored `oracle/_ref/` (listed in `.gitignore`, NOT in `.gpurunignore`, so it rides
along with the built `.so` files and never enters the history).  `__graft_entry__.build()` runs it whenever
`/root/reference` is present.  `oracle/ref_manifest.json` (committed: paths + SHA-256, no source) lets the GPU box prove
that what it runs is the unmodified reference.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs use the staged tree,
as the checker and as the timed baseline — never as part of the product path.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("B200DET_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
MANIFEST = os.path.join(HERE, "ref_manifest.json")

# the hot path (model/*.py::non_max_suppression, LightningFunc/accuracy.py, losses.py, step.py, utils/*) and what those
# modules import at module level; `dataset/pallete` is read by every model class body (e.g. model/YOLOV5.py:111)
TREES = ["LightningFunc", "model"]
FILES = ["dataset/pallete", "LICENSE"]


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def _listing(root):
    out = []
    for t in TREES:
        for dp, dn, fn in os.walk(os.path.join(root, t)):
            dn[:] = sorted(d for d in dn if d != "__pycache__")
            for f in sorted(fn):
                if f.endswith((".py", ".txt", ".yaml", ".yml", ".json")):
                    out.append(os.path.relpath(os.path.join(dp, f), root))
    out += [f for f in FILES if os.path.exists(os.path.join(root, f))]
    return out


def stage(write_manifest: bool = True) -> int:
    """Copy the reference files into oracle/_ref (verbatim) and write the manifest.  Returns the file count."""
    if not os.path.isdir(os.path.join(SRC, "LightningFunc")):
        return 0
    files = _listing(SRC)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    man = {}
    for rel in files:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
        os.chmod(dst, 0o644)
        man[rel] = _sha(dst)
    if write_manifest:
        with open(MANIFEST, "w") as f:
            json.dump({"source": "Leyan529/ObjectDetectionPL (MIT), staged verbatim from /root/reference",
                       "files": man}, f, indent=1, sort_keys=True)
            f.write("\n")
    return len(files)


def verify(root: str = DST):
    """Every staged file is byte-identical to what the manifest recorded from /root/reference.  Returns (ok, problems)."""
    if not os.path.exists(MANIFEST):
        return False, ["oracle/ref_manifest.json missing"]
    man = json.load(open(MANIFEST))["files"]
    bad = []
    for rel, sha in man.items():
        p = os.path.join(root, rel)
        if not os.path.exists(p):
            bad.append(f"missing {rel}")
        elif _sha(p) != sha:
            bad.append(f"modified {rel}")
    return not bad, bad


if __name__ == "__main__":
    n = stage()
    ok, bad = verify()
    print(f"staged {n} reference files into {DST}; manifest ok: {ok}", *bad, sep="\n")
    sys.exit(0 if ok or n == 0 else 1)

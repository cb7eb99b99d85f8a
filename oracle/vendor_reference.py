#!/usr/bin/env python
"""Places the reference's own hot-path modules under oracle/_ref/ so that the GPU box (which has no /root/reference) can
run the REAL reference model next to the drop-in.

TEST INFRASTRUCTURE ONLY.  oracle/_ref/ is git-ignored (the reference's sources never enter this repository's history)
but is not gpurun-ignored, so it travels with the snapshot like the built .so files.  Only tests/ may import it.

    python oracle/vendor_reference.py [/root/reference]

Copies, unmodified, the four files SURVEY.md 8(a) cites plus nothing else:
    src/liftreg/utils/sdct_projection_utils.py, src/liftreg/utils/net_utils.py, src/liftreg/layers/layers.py,
    src/liftreg/models/LiftRegDeformSubspaceBackproj.py
into oracle/_ref/liftreg/... (namespace packages: no __init__.py needed).  Returns the destination or None when the
reference tree is absent (the tests that need it then skip)."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ["utils/sdct_projection_utils.py", "utils/net_utils.py", "layers/layers.py", "models/LiftRegDeformSubspaceBackproj.py"]


def vendor(reference_root="/root/reference"):
    src = os.path.join(reference_root, "src", "liftreg")
    if not os.path.isdir(src):
        return None
    for rel in FILES:
        dst = os.path.join(DEST, "liftreg", rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), dst)
    return DEST


def available():
    return all(os.path.exists(os.path.join(DEST, "liftreg", rel)) for rel in FILES)


if __name__ == "__main__":
    out = vendor(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print(out if out else "reference tree not found; nothing vendored")

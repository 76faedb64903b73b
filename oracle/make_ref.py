"""TEST INFRASTRUCTURE ONLY -- copy recipe for oracle/_ref/ (git-ignored, travels to the GPU box with the snapshot).

The reference is pure Python: there is nothing to compile.  /root/reference does not exist on the GPU box, so the
integration tests that drive the reference's OWN `prepare_inputs_labels_for_multimodal`, `VideoProcessor` and
`max_coverage_sampling` loop next to the CUDA path (tests/test_gpu_overlay.py) need its files there.  This script
copies the Python sources of the `llava` package and the 3D preprocessing scripts, unmodified, from the mounted
reference into oracle/_ref/ -- outputs only, never committed (`oracle/_ref/` is listed in .gitignore).

    python oracle/make_ref.py            # also run by __graft_entry__.build() when /root/reference is mounted

oracle/ref_import.py imports from /root/reference when it is mounted and from oracle/_ref/ otherwise.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("V3D_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
TREES = ("llava", os.path.join("scripts", "3d", "preprocessing"))


def make(force: bool = False) -> str | None:
    if not os.path.isdir(os.path.join(SRC, "llava")):
        return None
    stamp = os.path.join(DST, ".complete")
    if os.path.isfile(stamp) and not force:
        return DST
    shutil.rmtree(DST, ignore_errors=True)
    n = 0
    for tree in TREES:
        for root, _dirs, files in os.walk(os.path.join(SRC, tree)):
            for fn in files:
                if not fn.endswith(".py"):
                    continue
                rel = os.path.relpath(os.path.join(root, fn), SRC)
                out = os.path.join(DST, rel)
                os.makedirs(os.path.dirname(out), exist_ok=True)
                shutil.copyfile(os.path.join(root, fn), out)
                n += 1
    with open(stamp, "w") as f:
        f.write(f"{n} files copied from {SRC}\n")
    return DST


if __name__ == "__main__":
    out = make(force="--force" in sys.argv)
    print(out if out else f"{SRC} is not mounted: nothing copied")

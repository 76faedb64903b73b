"""TEST INFRASTRUCTURE ONLY -- import the *unmodified* reference by file path.

Used in the build container (where /root/reference is mounted read-only) by
`oracle/gen_golden.py` to pin the oracle restatement and to write the golden
vectors under tests/golden/.  /root/reference does not exist on the GPU box: there the
integration tests (tests/test_gpu_overlay.py) import the unmodified copy that
`oracle/make_ref.py` placed under oracle/_ref/ (git-ignored); smoke() and bench.py never
import this module.

Recipe follows SURVEY.md section 8(c):
  * llava/video_utils.py and llava/model/position_encoding.py load standalone
    through importlib (their only deps are torch/numpy/cv2/PIL/tqdm).
  * llava/model/llava_arch.py needs a stub `llava` package (the real
    llava/__init__.py pulls the whole model zoo, which does not import under
    transformers 5.x) and a stub qformer module.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
# the mounted reference in the build container; on the GPU box the unmodified copy made by oracle/make_ref.py
REF_ROOT = os.environ.get("V3D_REFERENCE_ROOT") or ("/root/reference" if os.path.isdir("/root/reference/llava") else os.path.join(_HERE, "_ref"))


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "llava", "video_utils.py"))


def _load_by_path(name: str, relpath: str):
    path = os.path.join(REF_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_cache: dict = {}


def video_utils():
    """reference llava/video_utils.py (unproject, VideoProcessor, merge_video_dict)."""
    if "vu" not in _cache:
        _cache["vu"] = _load_by_path("_ref_video_utils", "llava/video_utils.py")
    return _cache["vu"]


def position_encoding():
    """reference llava/model/position_encoding.py."""
    if "pe" not in _cache:
        _cache["pe"] = _load_by_path("_ref_position_encoding", "llava/model/position_encoding.py")
    return _cache["pe"]


def llava_arch():
    """reference llava/model/llava_arch.py through a stub `llava` package."""
    if "arch" in _cache:
        return _cache["arch"]
    if "llava" not in sys.modules:
        pkg = types.ModuleType("llava")
        pkg.__path__ = [os.path.join(REF_ROOT, "llava")]
        sys.modules["llava"] = pkg
    qf_name = "llava.model.multimodal_resampler.qformer"
    if qf_name not in sys.modules:
        qf = types.ModuleType(qf_name)
        qf.Qformer = object
        sys.modules[qf_name] = qf
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):
        _cache["arch"] = importlib.import_module("llava.model.llava_arch")
    return _cache["arch"]


class ArchShim:
    """Minimal object exposing the four coordinate helpers of
    LlavaMetaForCausalLM (llava_arch.py:213-272) bound to a config namespace."""

    def __init__(self, voxel_size=0.1, min_xyz_range=(-15, -15, -5), max_xyz_range=(15, 15, 5)):
        arch = llava_arch()
        self.config = types.SimpleNamespace(
            voxel_size=voxel_size,
            min_xyz_range=list(min_xyz_range),
            max_xyz_range=list(max_xyz_range),
        )
        cls = arch.LlavaMetaForCausalLM
        for n in ("average_coordinate_in_patch", "minmax_coordinate_in_patch", "sample_n_points", "discrete_coords"):
            setattr(self, n, types.MethodType(getattr(cls, n), self))

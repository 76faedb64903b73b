"""Overlay launcher: run the reference's own scripts, unchanged, on top of the CUDA front end.

    torchrun ... -m video_3d_llm_b200.launch llava/train/train_3d.py <the script's own args>
    python -m video_3d_llm_b200.launch llava/eval/model_scanqa.py --frame_sampling_strategy mc-ratio95 ...
    python -m video_3d_llm_b200.launch scripts/3d/preprocessing/max_coverage_sampling.py --voxel_size 0.1

`install()` rebinds, before the script is executed,
  llava.video_utils.{unproject, VideoProcessor, merge_video_dict}          (bound by `from ... import` at
      script import: train_3d.py:48, model_scanqa.py:18, max_coverage_sampling.py:6)
  llava.model.position_encoding.{PositionEmbeddingSine3D, PositionEmbeddingMLP}   (imported lazily inside
      LlavaMetaModel.__init__, llava_arch.py:49)
  LlavaMetaForCausalLM.{average_coordinate_in_patch, minmax_coordinate_in_patch, sample_n_points,
      discrete_coords}                                                      (called at llava_arch.py:396-420)
  LlavaMetaForCausalLM.{get_2dPool, add_token_per_grid}                     (called at :469 and :536; the slow-fast
      variant `add_faster_video`, which this package does not implement, keeps the reference's own method)
The 500-line prepare_inputs_labels_for_multimodal itself is left alone: with VideoProcessor(lazy=True) the
`world_coords` it receives is a handle over raw depth/pose/intrinsics and the patched helpers return deferred values
(lazy.py), so that the method's own call sequence ends in fused launches:
  avg-discrete-sin3d video sample:  v3d_frontend_pe with the PE off (depth -> discretised patch coordinates, 9 KB per frame)
                                    + v3d_pool_pe_grid (get_2dPool + PE + add + newline over the projector's tokens),
  without spatial pooling / newline: ONE v3d_frontend_pe launch with tokens_in (depth + tokens -> tokens + PE),
and the (V,384,384,3) coordinate tensor, the pooled tokens and the stand-alone PE tensor are never built.
"""
from __future__ import annotations

import os
import runpy
import sys


_originals: list = []  # (object, attribute name, original value) of everything install() rebound


def uninstall() -> None:
    """Restore what install() rebound (tests run the un-patched reference next to the patched one)."""
    from . import position_encoding

    while _originals:
        obj, name, val = _originals.pop()
        setattr(obj, name, val)
    position_encoding.PositionEmbeddingSine3D.defer_add = False


def _rebind(obj, name, val) -> None:
    _originals.append((obj, name, getattr(obj, name)))
    setattr(obj, name, val)


def install(lazy: bool | None = None) -> dict:
    """Patch the reference modules in place; returns {dotted name: replacement} for inspection."""
    import importlib

    uninstall()

    from . import cabi, coords, position_encoding, projector_tail, video_utils
    from . import lazy as lazy_mod

    if lazy is None:
        lazy = os.environ.get("V3D_LAZY", "1") != "0"
    # V3D_DEFER=0 keeps every patched helper eager (one launch per call, real tensors everywhere)
    defer = os.environ.get("V3D_DEFER", "1") != "0"
    position_encoding.PositionEmbeddingSine3D.defer_add = defer
    done = {}

    vu = importlib.import_module("llava.video_utils")

    class VideoProcessor(video_utils.VideoProcessor):
        def __init__(self, *a, **kw):
            kw.setdefault("lazy", lazy)
            if os.environ.get("V3D_PACK_DIR"):
                kw.setdefault("pack_dir", os.environ["V3D_PACK_DIR"])
            super().__init__(*a, **kw)
            if os.environ.get("V3D_MC_PRECOMPUTE", "0") == "1":
                # 'mc-online*' strategies: DataLoader workers are forked and cannot touch CUDA -- sample every scene here
                self.precompute_mc_order()

    for name, obj in (("unproject", video_utils.unproject), ("VideoProcessor", VideoProcessor),
                      ("merge_video_dict", video_utils.merge_video_dict)):
        _rebind(vu, name, obj)
        done[f"llava.video_utils.{name}"] = obj

    pe = importlib.import_module("llava.model.position_encoding")
    for name in ("PositionEmbeddingSine3D", "PositionEmbeddingMLP"):
        _rebind(pe, name, getattr(position_encoding, name))
        done[f"llava.model.position_encoding.{name}"] = getattr(position_encoding, name)

    arch = importlib.import_module("llava.model.llava_arch")
    cls = arch.LlavaMetaForCausalLM

    def average_coordinate_in_patch(self, world_coords, patch_size=27):
        return coords.average_coordinate_in_patch(world_coords, patch_size)

    def minmax_coordinate_in_patch(self, world_coords, patch_size=27):
        return coords.minmax_coordinate_in_patch(world_coords, patch_size)

    def sample_n_points(self, world_coords, n_points=9):
        return coords.sample_n_points(world_coords, n_points)

    def discrete_coords(self, world_coords, xyz_min):
        # the reference runs this on CUDA, where ATen multiplies by the reciprocal of the scalar divisor (cabi.div_mode_default)
        return coords.discrete_coords(world_coords, xyz_min, voxel_size=self.config.voxel_size,
                                      min_xyz_range=self.config.min_xyz_range, max_xyz_range=self.config.max_xyz_range,
                                      div_mode=cabi.div_mode_default())

    ref_get_2dpool, ref_add_token_per_grid = cls.get_2dPool, cls.add_token_per_grid

    def get_2dPool(self, image_feature, stride=2):
        side = self.get_vision_tower().num_patches_per_side
        mode = self.config.mm_spatial_pool_mode
        if side * side != image_feature.shape[1]:
            return ref_get_2dpool(self, image_feature, stride)
        cabi.require_cuda(image_feature, "image_feature")
        if defer and mode in ("bilinear", "average", "max"):
            # deferred: `+ PE` (:515) and add_token_per_grid (:536) complete it in one kernel over the un-pooled tokens
            return lazy_mod.LazyTokenPool(image_feature, stride, mode)
        return projector_tail.get_2dPool(image_feature, stride, mode)

    def add_token_per_grid(self, image_feature):
        if getattr(self.config, "add_faster_video", False):
            return ref_add_token_per_grid(self, lazy_mod._unwrap(image_feature))
        if isinstance(image_feature, lazy_mod.LazyTokenPool):
            return image_feature.finish(self.model.image_newline)
        return projector_tail.add_token_per_grid(lazy_mod._unwrap(image_feature), self.model.image_newline)

    for fn in (average_coordinate_in_patch, minmax_coordinate_in_patch, sample_n_points, discrete_coords, get_2dPool, add_token_per_grid):
        _rebind(cls, fn.__name__, fn)
        done[f"llava.model.llava_arch.LlavaMetaForCausalLM.{fn.__name__}"] = fn
    return done


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit(__doc__)
    script = argv[0]
    if os.path.basename(script) == "max_coverage_sampling.py":
        # the offline MC script is replaced wholesale: same CLI, same JSON (mc_sampling.main)
        from . import mc_sampling

        sys.argv = argv
        runpy.run_module(mc_sampling.__name__, run_name="__main__")
        return
    install()
    sys.argv = argv
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()

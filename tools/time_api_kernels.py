"""Timing of the API-parity kernels at the bench's batch (512 frames of 480x640): v3d_unproject (full frames and the
resampled 384x384 window), v3d_boundary, v3d_pool_coords + v3d_discretize.  python tools/time_api_kernels.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__  # noqa: F401
from video_3d_llm_b200 import coords as C, geometry, synthetic, video_utils as VU

dev = torch.device("cuda:0")
scenes = [synthetic.make_scene(100 + s, 32, n_total=300, device=dev) for s in range(16)]
depth = torch.cat([s.depth for s in scenes]); poses = torch.cat([s.poses for s in scenes]); intr = torch.cat([s.intrinsics for s in scenes])
rm, cm = geometry.resample_maps(480, 640, 384, "center_crop", dev)


def timed(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


V = depth.shape[0]
res = {}
ms = timed(lambda: VU.unproject(intr, poses, depth))
res["unproject_full"] = {"ms": ms, "GBs": V * (480 * 640 * 14 + 80) / ms / 1e6}
ms = timed(lambda: VU.unproject(intr, poses, depth, row_map=rm, col_map=cm))
res["unproject_384_window"] = {"ms": ms, "GBs": V * (384 * 384 * 14 + 80) / ms / 1e6}
ms = timed(lambda: VU.boundary(intr, poses, depth))
res["boundary"] = {"ms": ms, "GBs": V * (480 * 640 * 2 + 80) / ms / 1e6}
c384 = VU.unproject(intr[:128], poses[:128], depth[:128], row_map=rm, col_map=cm)
ms = timed(lambda: C.average_coordinate_in_patch(c384, 27))
res["pool_coords_avg27_128frames"] = {"ms": ms, "GBs": 128 * 378 * 378 * 12 / ms / 1e6}
print(res)

# object proposals: 32 frames, 100 boxes
sc = scenes[0]
boxes = torch.rand(100, 6, device=dev) * torch.tensor([4.0, 4.0, 2.0, 1.5, 1.5, 1.5], device=dev) + torch.tensor([-2.0, -2.0, 0.0, 0.2, 0.2, 0.2], device=dev)
c32 = VU.unproject(sc.intrinsics, sc.poses, sc.depth, row_map=rm, col_map=cm)
lazy = VU.LazyWorldCoords(sc.depth, sc.poses, sc.intrinsics)
obj = {}
for kind in ("patch14", "patch27"):
    obj[kind + "_from_coords_ms"] = timed(lambda: C.object_patch_masks(c32, boxes, kind))
    obj[kind + "_from_depth_ms"] = timed(lambda: C.object_patch_masks(lazy, boxes, kind))
masks = C.object_patch_masks(c32, boxes, "patch27")
feats = torch.randn(32, 196, 3584, device=dev).to(torch.bfloat16)
obj["masked_token_mean_ms"] = timed(lambda: C.object_features(feats, masks))
print({"object_proposals_32_frames_100_boxes": obj})

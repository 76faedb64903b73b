"""Time the MC sampler stages on the bench's MC workload (16 scenes x 150 candidate frames, 480x640).

    python tools/time_mc.py [reps] [n_scenes] [candidates: 150 (every second frame of 300) or 300 (all frames)]

CUDA events on the launching stream, best and median of `reps` runs per stage after one warm-up.
V3D_MC_DENSE=0 in the environment times the hash voxeliser alone (k_mc_voxelize without k_mc_voxelize_dense).
"""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import __graft_entry__  # noqa: F401  (package alias)
from video_3d_llm_b200 import cabi, synthetic
from video_3d_llm_b200.mc_sampling import MCSampler

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
n_scenes = int(sys.argv[2]) if len(sys.argv) > 2 else 16
cand = int(sys.argv[3]) if len(sys.argv) > 3 else 150
dev = torch.device("cuda:0")
scenes = []
for s in range(n_scenes):
    seed = 500 + s
    sc = synthetic.make_scene(seed, cand, n_total=300, frame_indices=np.arange(0, 300, 300 // cand), device=dev)
    pc = torch.from_numpy(synthetic.mesh_voxels(sc.room, sc.boxes, seed)).to(dev)
    scenes.append({"depth": sc.depth, "poses": sc.poses, "intrinsics": sc.intrinsics, "pc_voxels": pc, "seed": seed})
sampler = MCSampler(device=dev)
plan = sampler.plan(scenes)
sampler.run(plan)
torch.cuda.synchronize()
out = {"lib": cabi.LIB_PATH, "n_scenes": n_scenes, "candidates": cand, "n_pc_mean": float(np.mean([s["pc_voxels"].shape[0] for s in scenes]))}
for stage in ("voxelize", "greedy", "all"):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sampler.run(plan, stage)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    out[stage + "_ms_best"] = min(ts)
    out[stage + "_ms_median"] = statistics.median(ts)
px = n_scenes * cand * 480 * 640
out["voxelize_gpx_per_s"] = px / out["voxelize_ms_best"] / 1e6
out["voxelize_hbm_gbs"] = px * 2 / out["voxelize_ms_best"] / 1e6
out["scenes_per_s"] = n_scenes / out["all_ms_median"] * 1e3
print(out)

"""Time k_fused at BASELINE config 2 (27x27 grid, E=1152) and at the default recipe (14x14, E=3584), fused add.
python tools/time_fused.py [reps]   (V3D_LIB_PATH selects an experimental build)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__  # noqa: F401
from video_3d_llm_b200 import cabi, synthetic
from video_3d_llm_b200.frontend import FrontEnd

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda:0")
scenes = [synthetic.make_scene(100 + s, 32, n_total=300, device=dev) for s in range(16)]
depth = torch.cat([s.depth for s in scenes])
poses = torch.cat([s.poses for s in scenes])
intr = torch.cat([s.intrinsics for s in scenes])
res = {"lib": os.path.basename(cabi.LIB_PATH)}
for name, patch, E, G in (("config2_27x27_E1152", 14, 1152, 27), ("default_14x14_E3584", 27, 3584, 14)):
    fe = FrontEnd(device=dev, patch_size=patch, embedding_size=E)
    tok = torch.randn((512, G * G, E), device=dev).to(torch.bfloat16)
    out = torch.empty_like(tok)
    for _ in range(5):
        fe(depth, poses, intr, tokens=tok, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fe(depth, poses, intr, tokens=tok, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    byts = (285848 + 2 * G * G * E * 2) * 512
    res[name] = {"ms": round(ms, 4), "GBs": round(byts / ms / 1e6, 1), "frac_of_6545.6": round(byts / ms / 1e6 / 6545.6, 4)}
    del tok, out
print(res)

"""Default recipe (14x14 grid, E=3584): fused kernel vs the split at the pooled-coordinates boundary
(kernel A: depth -> discretised patch coordinates; kernel B: PE + token add)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__  # noqa: F401
from video_3d_llm_b200 import projector_tail as PT, synthetic
from video_3d_llm_b200.frontend import FrontEnd
from video_3d_llm_b200.position_encoding import PositionEmbeddingSine3D

dev = torch.device("cuda:0")
scenes = [synthetic.make_scene(100 + s, 32, n_total=300, device=dev) for s in range(16)]
depth = torch.cat([s.depth for s in scenes]); poses = torch.cat([s.poses for s in scenes]); intr = torch.cat([s.intrinsics for s in scenes])
E, G = 3584, 14
tok = torch.randn((512, G * G, E), device=dev).to(torch.bfloat16)
out = torch.empty_like(tok)
fe = FrontEnd(device=dev, patch_size=27, embedding_size=E)
fa = FrontEnd(device=dev, patch_size=27, embedding_size=E, world_position_embedding_type="avg-discrete")
pe = PositionEmbeddingSine3D(E)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


_, _, disc = fa(depth, poses, intr, want_coords=True)
disc = disc.reshape(512, G * G, 3)
res = {"fused_ms": timed(lambda: fe(depth, poses, intr, tokens=tok, out=out)),
       "A_coords_only_ms": timed(lambda: fa(depth, poses, intr, want_coords=True)),
       "B_pe_module_add_ms": timed(lambda: pe(disc.to(torch.bfloat16), add_to=tok)),
       "B_tail_kernel_none_ms": timed(lambda: PT.pool_pe_newline(tok, mode="none", coords=disc, position_embedding=pe))}
print(res)

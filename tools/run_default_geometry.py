"""One process running the fused kernel at the reference-default geometry (patch 27 -> 14x14, E=3584): ncu target."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_3d_llm_b200 import synthetic
from video_3d_llm_b200.frontend import FrontEnd
dev = torch.device("cuda:0")
scenes = [synthetic.make_scene(100 + s, 32, n_total=300, device=dev) for s in range(16)]
depth = torch.cat([s.depth for s in scenes]); poses = torch.cat([s.poses for s in scenes]); intr = torch.cat([s.intrinsics for s in scenes])
fe = FrontEnd(device=dev, patch_size=27, embedding_size=3584)
tok = torch.randn((512, 196, 3584), device=dev).to(torch.bfloat16); out = torch.empty_like(tok)
for _ in range(5): fe(depth, poses, intr, tokens=tok, out=out)
torch.cuda.synchronize(); print("ok")

"""One process running the fused front end a few times (ncu target): python tools/run_fused_once.py [patch] [E] [exact]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_3d_llm_b200 import synthetic
from video_3d_llm_b200.frontend import FrontEnd
patch = int(sys.argv[1]) if len(sys.argv) > 1 else 14
E = int(sys.argv[2]) if len(sys.argv) > 2 else 1152
exact = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda:0")
scenes = [synthetic.make_scene(100 + s, 32, n_total=300, device=dev) for s in range(16)]
depth = torch.cat([s.depth for s in scenes]); poses = torch.cat([s.poses for s in scenes]); intr = torch.cat([s.intrinsics for s in scenes])
fe = FrontEnd(device=dev, patch_size=patch, embedding_size=E)
fe.exact_pooling = exact
G = 378 // patch
tok = torch.randn((512, G * G, E), device=dev).to(torch.bfloat16); out = torch.empty_like(tok)
for _ in range(4): fe(depth, poses, intr, tokens=tok, out=out)
torch.cuda.synchronize(); print("ok")

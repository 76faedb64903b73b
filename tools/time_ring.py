"""Time the fused front end (ring kernel) for both geometries: python tools/time_ring.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_3d_llm_b200 import synthetic
from video_3d_llm_b200.frontend import FrontEnd
dev = torch.device('cuda:0')
scenes = [synthetic.make_scene(100 + s, 32, n_total=300, device=dev) for s in range(16)]
depth = torch.cat([s.depth for s in scenes]); poses = torch.cat([s.poses for s in scenes]); intr = torch.cat([s.intrinsics for s in scenes])
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for patch, E in ((14, 1152), (27, 3584)):
    G = 378 // patch
    tok = torch.randn((512, G * G, E), device=dev).to(torch.bfloat16); out = torch.empty_like(tok)
    fe = FrontEnd(device=dev, patch_size=patch, embedding_size=E)
    ms = min(timed(lambda: fe(depth, poses, intr, tokens=tok, out=out)) for _ in range(3))
    byts = (378*378*2+80 + 2*G*G*E*2) * 512
    print(f"{os.environ.get('V3D_RING_POOL','')} patch {patch} E {E}: {ms*1e3:.1f} us = {byts/ms/1e6:.0f} GB/s ({byts/ms/1e6/6545.6:.3f})")
    # one 32-frame scene per launch (BASELINE configs[0]), rotating over the 16 scenes so that inputs come from HBM
    k = [0]
    def one():
        b = k[0] % 16; k[0] += 1
        a, e = b * 32, (b + 1) * 32
        fe(depth[a:e], poses[a:e], intr[a:e], tokens=tok[a:e], out=out[a:e])
    ms1 = min(timed(one, reps=64) for _ in range(3))
    print(f"   one scene per launch: {ms1*1e3:.1f} us = {byts/16/ms1/1e6:.0f} GB/s ({byts/16/ms1/1e6/6545.6:.3f})")
    del tok, out

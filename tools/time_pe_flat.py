"""Stand-alone PE (+ add) on discretised coordinates (PositionEmbeddingSine3D.forward, k_pe_flat): python tools/time_pe_flat.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_3d_llm_b200.position_encoding import PositionEmbeddingSine3D
dev = torch.device("cuda:0")
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for E, n_tok in ((1152, 512 * 729), (3584, 512 * 196)):
    pe = PositionEmbeddingSine3D(E)
    c = torch.randint(0, 101, (1, n_tok, 3), device=dev).to(torch.bfloat16)
    tok = torch.randn((1, n_tok, E), device=dev).to(torch.bfloat16)
    ms = min(timed(lambda: pe(c, add_to=tok)) for _ in range(3))
    byts = 2 * n_tok * E * 2 + n_tok * 12
    ms1 = min(timed(lambda: pe(c)) for _ in range(3))
    print(f"E {E} tokens {n_tok}: PE + add {ms*1e3:.1f} us = {byts/ms/1e6:.0f} GB/s ({byts/ms/1e6/6545.6:.3f}); PE only {ms1*1e3:.1f} us = {(n_tok*E*2)/ms1/1e6:.0f} GB/s")

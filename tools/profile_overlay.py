"""Host cost of the overlay's per-scene call sequence (helpers -> deferred handles -> one fused launch): python tools/profile_overlay.py"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from video_3d_llm_b200 import synthetic, video_utils as VU
from video_3d_llm_b200.position_encoding import PositionEmbeddingSine3D
dev = torch.device("cuda:0")
sc = synthetic.make_scene(3, 32, n_total=300, device=dev)
h = VU.LazyWorldCoords(sc.depth, sc.poses, sc.intrinsics, src_hw=(480, 640))
tok = torch.randn((32, 729, 1152), device=dev).to(torch.bfloat16)
pe = PositionEmbeddingSine3D(1152); pe.defer_add = True
f = lambda: bench.reference_call_sequence(h, tok, pe, 14)
for _ in range(20): f()
torch.cuda.synchronize()
N = 300
t0 = time.perf_counter()
for _ in range(N): f()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host issue time per scene: {(t1 - t0) / N * 1e6:.1f} us; with the GPU drained: {(t2 - t0) / N * 1e6:.1f} us")
pr = cProfile.Profile(); pr.enable()
for _ in range(N): f()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)

"""Small invocations of every kernel family, for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
from video_3d_llm_b200 import synthetic, coords as C, video_utils as VU
from video_3d_llm_b200.frontend import FrontEnd
from video_3d_llm_b200.position_encoding import PositionEmbeddingSine3D, PositionEmbeddingMLP
dev = torch.device("cuda:0")
g.smoke()
sc = synthetic.make_scene(5, 2, n_total=9, device="cpu")
d, p, k = sc.depth.to(dev), sc.poses.to(dev), sc.intrinsics.to(dev)
for kind, patch, E in (("avg-discrete-sin3d", 27, 3584), ("minmax-discrete-sin3d", 27, 3584), ("sample9-discrete-sin3d", 27, 3584), ("avg-sin3d", 14, 1152)):
    fe = FrontEnd(device=dev, patch_size=patch, embedding_size=E, world_position_embedding_type=kind)
    tok = synthetic.make_tokens(1, 2, fe.tokens_per_frame, E).to(dev)
    fe(d, p, k, tokens=tok, want_coords=True)
w = VU.unproject(k, p, d); VU.boundary(k, p, d)
from video_3d_llm_b200 import geometry
rm, cm = geometry.resample_maps(480, 640, 384, "center_crop", dev)
c384 = VU.unproject(k, p, d, row_map=rm, col_map=cm)
C.average_coordinate_in_patch(c384, 27); C.minmax_coordinate_in_patch(c384, 14); C.sample_n_points(c384, 5); C.discrete_coords(c384[:, :10, :10])
PositionEmbeddingSine3D(1155)(torch.rand(1, 7, 3, device=dev) * 30)
boxes = torch.tensor([[0, 0, 1, 2, 2, 2], [1, 1, 1, 0.5, 0.5, 0.5]], dtype=torch.float32)
m = C.object_patch_masks(c384, boxes, "patch14"); C.object_features(synthetic.make_tokens(2, 2, 729, 64).to(dev), m)
mod = PositionEmbeddingMLP(192).to(dev).eval()
with torch.no_grad():
    mod(torch.rand(1, 300, 3, device=dev), add_to=torch.randn(1, 300, 192, device=dev).to(torch.bfloat16))
torch.cuda.synchronize(); print("sanitize run ok")

"""Stage-by-stage comparison of the reference arm and the overlay arm of tests/test_gpu_overlay.py (GPU box)."""
import importlib, json, os, sys, tempfile
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fixtures, ref_harness
from oracle import ref_import
from video_3d_llm_b200 import launch, coords as C, video_utils as VU
DEV = torch.device("cuda:0")
D = json.load(open(os.path.join(ROOT, "tests/golden/dataset.json")))
root = tempfile.mkdtemp()
m = D["scenes"]["scene0003_00"]
fixtures.write_dataset(root, [(m["seed"], m["n_total"], m["split"])], H=480, W=640)
os.chdir(root)
dtype = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[sys.argv[1] if len(sys.argv) > 1 else "f32"]
arch = ref_import.llava_arch(); rvu = ref_import.video_utils(); rpe = ref_import.position_encoding()
kw = dict(video_folder="data", annotation_dir="data/embodiedscan", voxel_size=0.1, min_xyz_range=[-15,-15,-5], max_xyz_range=[15,15,5], frame_sampling_strategy="uniform")
vd = rvu.VideoProcessor(**kw).process_3d_video("scannet/scene0003_00", fixtures.FakeImageProcessor(384), force_sample=True, frames_upbound=4)
wc = vd["world_coords"].to(DEV).to(dtype)
shim = ref_import.ArchShim()
vp = VU.VideoProcessor(**kw, lazy=True)
vd2 = vp.process_3d_video("scannet/scene0003_00", fixtures.FakeImageProcessor(384), force_sample=True, frames_upbound=4)
h = vd2["world_coords"].to(dtype).to(DEV)
def cmp(name, a, b):
    a, b = a.float(), b.float()
    d = (a - b).abs()
    print(f"{name}: shape {tuple(a.shape)} max|d| {d.max().item():.3e} frac!=0 {(d > 0).float().mean().item():.3e} max|ref| {a.abs().max().item():.3e}")
print("materialised vs reference world coords:"); cmp("world", wc, h.materialize())
p_ref = shim.average_coordinate_in_patch(wc); p_our = C.average_coordinate_in_patch(h).materialize()
cmp("pooled", p_ref, p_our)
d_ref = shim.discrete_coords(p_ref, None); d_our = C.discrete_coords(C.average_coordinate_in_patch(h), None, div_mode=1).materialize()
cmp("disc(recip)", d_ref, d_our)
d_our0 = C.discrete_coords(C.average_coordinate_in_patch(h), None, div_mode=0).materialize()
cmp("disc(ieee)", d_ref, d_our0)
E = 256
pe_ref = rpe.PositionEmbeddingSine3D(E)(d_ref.flatten(1, 2))
from video_3d_llm_b200.position_encoding import PositionEmbeddingSine3D
pe_our = PositionEmbeddingSine3D(E)(d_ref.flatten(1, 2))
cmp("pe(same coords)", pe_ref, pe_our)
tok = torch.randn(4, 729, E, device=DEV).to(dtype)
from video_3d_llm_b200 import projector_tail as PT
x = tok.view(4, 27, 27, -1).permute(0, 3, 1, 2).contiguous()
x = torch.nn.functional.interpolate(x, size=[14, 14], mode="bilinear").permute(0, 2, 3, 1).reshape(4, 196, E)
cmp("pool", x, PT.get_2dPool(tok, 2, "bilinear"))
cmp("pool+pe", x + pe_ref, PT.pool_pe_newline(tok, coords=d_ref.flatten(1, 2), position_embedding=PositionEmbeddingSine3D(E), flatten=False))

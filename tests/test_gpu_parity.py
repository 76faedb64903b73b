"""GPU: parity of the CUDA path (through the host mirror -> ctypes -> C ABI) against the CPU oracle
on the same seeded inputs, and against the golden vectors produced by the reference.

Bar (BASELINE.json north_star): integer / index work and fp32 coordinates bit-exact; PE within bf16
rounding of the reference's fp32 value (sin/cos come from different math libraries, so an element
may land on the other side of a bf16 tie: we bound the fraction of such elements and their size).
"""
import json
import os

import numpy as np
import pytest
import torch

import fixtures
from oracle import oracle as O

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

from video_3d_llm_b200 import cabi, coords as C, mc_sampling, synthetic, video_utils as VU  # noqa: E402
from video_3d_llm_b200.frontend import FrontEnd  # noqa: E402
from video_3d_llm_b200.position_encoding import PositionEmbeddingSine3D  # noqa: E402

DEV = torch.device("cuda:0")
BF16_ULP_TOL = 1  # at most one bf16 ulp on the rare element that flips


def bf16_bits(t: torch.Tensor) -> np.ndarray:
    return t.contiguous().view(torch.int16).cpu().numpy().view(np.uint16)


def assert_bf16_close(got_bits: np.ndarray, ref_bits: np.ndarray, max_frac=2e-4):
    """Equal except for at most max_frac of the elements, which may differ by one bf16 ulp."""
    assert got_bits.shape == ref_bits.shape
    diff = got_bits != ref_bits
    frac = float(diff.mean())
    assert frac <= max_frac, f"{frac:.3e} of the bf16 elements differ"
    if diff.any():
        g = got_bits[diff].astype(np.int32)
        r = ref_bits[diff].astype(np.int32)
        # same sign: adjacent bit patterns; across zero (+-tiny) compare values instead
        same_sign = (g >> 15) == (r >> 15)
        assert np.all(np.abs(g[same_sign] - r[same_sign]) <= BF16_ULP_TOL)
        gv, rv = O.bf16_bits_to_f32(got_bits[diff][~same_sign]), O.bf16_bits_to_f32(ref_bits[diff][~same_sign])
        assert np.all(np.abs(gv - rv) < 1e-6)


@pytest.fixture(scope="module")
def scene():
    return synthetic.make_scene(11, 3, n_total=40, device="cpu")


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "frontend.npz"))


def _dev(sc):
    return sc.depth.to(DEV), sc.poses.to(DEV), sc.intrinsics.to(DEV)


# ---------------------------------------------------------------------------------------------
# unprojection
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", ["u16", "i32", "f32"])
def test_unproject_bit_exact(scene, dtype):
    d, p, k = _dev(scene)
    du = scene.depth_u16()
    if dtype == "i32":
        d = torch.from_numpy(du.astype(np.int32)).to(DEV)
    elif dtype == "f32":
        d = torch.from_numpy(du.astype(np.float32)).to(DEV)
    got = VU.unproject(k, p, d).cpu().numpy()
    ref = O.unproject(scene.intrinsics.numpy(), scene.poses.numpy(), du)
    np.testing.assert_array_equal(got, ref)


def test_unproject_golden(G):
    sc = synthetic.make_scene(0, 2, n_total=30, device="cpu")
    got = VU.unproject(sc.intrinsics.to(DEV), sc.poses.to(DEV), sc.depth.to(DEV)).cpu().numpy()
    np.testing.assert_array_equal(got[:, ::37, ::41], G["world_sample"])
    import hashlib
    assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest() == str(G["world_sha"])


def test_unproject_accepts_cpu_inputs_and_odd_shapes():
    sc = synthetic.make_scene(3, 2, n_total=9, H=37, W=53, device="cpu")  # W*2 not a multiple of 16, sizes not /4
    got = VU.unproject(sc.intrinsics, sc.poses, sc.depth)
    assert got.is_cuda and got.shape == (2, 37, 53, 3)
    np.testing.assert_array_equal(got.cpu().numpy(), O.unproject(sc.intrinsics.numpy(), sc.poses.numpy(), sc.depth_u16()))


def test_unproject_non_rigid_pose_and_nonfinite(scene):
    d, p, k = _dev(scene)
    p = p.clone()
    p[0, 3] = torch.tensor([0.01, -0.02, 0.03, 1.5], device=DEV)  # projective bottom row -> true division by w
    p[1, 0, 3] = float("inf")
    got = VU.unproject(k, p, d).cpu().numpy()
    ref = O.unproject(scene.intrinsics.numpy(), p.cpu().numpy(), scene.depth_u16())
    np.testing.assert_array_equal(got, ref)  # NaN/Inf compare equal position-wise in assert_array_equal


def test_unproject_resampled_and_clamped(scene):
    d, p, k = _dev(scene)
    from video_3d_llm_b200 import geometry
    rm, cm = geometry.resample_maps(480, 640, 384, "center_crop", DEV)
    lo, hi = [-1.0, -1.5, 0.2], [1.2, 1.0, 2.0]
    got = VU.unproject(k, p, d, row_map=rm, col_map=cm, clamp=(lo, hi)).cpu().numpy()
    world = O.clamp(O.unproject(scene.intrinsics.numpy(), scene.poses.numpy(), scene.depth_u16()), lo, hi)
    orm, ocm = O.resample_maps(480, 640)
    np.testing.assert_array_equal(got, O.resample(world, orm, ocm))


def test_boundary(scene):
    d, p, k = _dev(scene)
    got = VU.boundary(k, p, d).cpu().numpy()
    ref = O.boundary(O.unproject(scene.intrinsics.numpy(), scene.poses.numpy(), scene.depth_u16()))
    np.testing.assert_array_equal(got, ref)


# ---------------------------------------------------------------------------------------------
# helpers on materialised coordinates
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def coords384(scene):
    world = O.unproject(scene.intrinsics.numpy(), scene.poses.numpy(), scene.depth_u16())
    rm, cm = O.resample_maps(480, 640)
    return O.resample(world, rm, cm)


@pytest.mark.parametrize("p", [27, 14])
def test_pool_helpers_bit_exact(coords384, p):
    ct = torch.from_numpy(coords384).to(DEV)
    np.testing.assert_array_equal(C.average_coordinate_in_patch(ct, p).cpu().numpy(), O.average_coordinate_in_patch(coords384, p))
    np.testing.assert_array_equal(C.minmax_coordinate_in_patch(ct, p).cpu().numpy(), O.minmax_coordinate_in_patch(coords384, p))


@pytest.mark.parametrize("n", [9, 5, 1])
def test_sample_n_points(coords384, n):
    ct = torch.from_numpy(coords384).to(DEV)
    np.testing.assert_array_equal(C.sample_n_points(ct, n).cpu().numpy(), O.sample_n_points(coords384, n))


def test_pool_helpers_golden(G):
    sc = synthetic.make_scene(0, 2, n_total=30, device="cpu")
    from video_3d_llm_b200 import geometry
    rm, cm = geometry.resample_maps(480, 640, 384, "center_crop", DEV)
    ct = VU.unproject(sc.intrinsics, sc.poses, sc.depth, row_map=rm, col_map=cm)
    np.testing.assert_array_equal(C.average_coordinate_in_patch(ct, 27).cpu().numpy(), G["avg27"])
    np.testing.assert_array_equal(C.average_coordinate_in_patch(ct, 14).cpu().numpy(), G["avg14"])
    np.testing.assert_array_equal(C.minmax_coordinate_in_patch(ct, 27).cpu().numpy(), G["minmax27"])
    np.testing.assert_array_equal(C.sample_n_points(ct, 9).cpu().numpy(), G["sample9"])
    np.testing.assert_array_equal(C.discrete_coords(torch.from_numpy(G["avg27"]).to(DEV)).cpu().numpy(), G["disc27"])


def test_pool_nan_propagation():
    c = np.random.default_rng(0).standard_normal((1, 384, 384, 3)).astype(np.float32)
    c[0, 30, 40, 1] = np.nan
    c[0, 100, 200, 2] = np.inf
    ct = torch.from_numpy(c).to(DEV)
    np.testing.assert_array_equal(C.average_coordinate_in_patch(ct, 27).cpu().numpy(), O.average_coordinate_in_patch(c, 27))
    np.testing.assert_array_equal(C.minmax_coordinate_in_patch(ct, 27).cpu().numpy(), O.minmax_coordinate_in_patch(c, 27))


def test_discretize(G):
    x = np.random.default_rng(1).uniform(-20, 20, (5000, 3)).astype(np.float32)
    x[:6] = G["disc_stress_in"]
    xt = torch.from_numpy(x).to(DEV)
    np.testing.assert_array_equal(C.discrete_coords(xt).cpu().numpy(), O.discrete_coords(x))
    np.testing.assert_array_equal(C.discrete_coords(xt[:6]).cpu().numpy(), G["disc_stress_out"])
    np.testing.assert_array_equal(C.discrete_coords(xt, div_mode=cabi.DIV_RECIPROCAL).cpu().numpy(),
                                  O.discrete_coords(x, div_mode=O.DIV_RECIPROCAL))
    got = C.discrete_coords(xt, min_xyz_range=None, max_xyz_range=None).cpu().numpy()
    np.testing.assert_array_equal(got, np.rint(x / np.float32(0.1)))
    # dtype follows the input like the reference
    assert C.discrete_coords(xt.half()).dtype == torch.float16


# ---------------------------------------------------------------------------------------------
# sin-3D PE
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("E,n_points", [(3584, 1), (1152, 1), (1155, 1), (3584, 9), (3584, 2), (96, 1)])
def test_pe_module_discrete_inputs(E, n_points):
    rng = np.random.default_rng(E + n_points)
    shape = (2, 50, 3) if n_points == 1 else (2, 50, n_points, 3)
    x = np.stack([rng.integers(0, 301, shape[:-1]), rng.integers(0, 301, shape[:-1]), rng.integers(0, 101, shape[:-1])], -1).astype(np.float32)
    mod = PositionEmbeddingSine3D(E, n_points=n_points)
    ref = O.pe_sine3d(x, E, n_points=n_points)
    got32 = mod(torch.from_numpy(x).to(DEV))
    assert got32.dtype == torch.float32 and got32.shape == ref.shape
    np.testing.assert_allclose(got32.cpu().numpy(), ref, rtol=0, atol=5e-7)
    gotbf = mod(torch.from_numpy(x).to(DEV).to(torch.bfloat16))  # bf16 inputs: integers above 256 are not representable
    xb = torch.from_numpy(x).to(torch.bfloat16).float().numpy()
    assert_bf16_close(bf16_bits(gotbf), O.f32_to_bf16_bits(O.pe_sine3d(xb, E, n_points=n_points)))
    used = 3 * n_points * (E // (3 * n_points))
    assert torch.all(got32[..., used:] == 0)


def test_pe_module_continuous_and_special_inputs():
    rng = np.random.default_rng(5)
    x = rng.uniform(-15, 15, (2, 64, 3)).astype(np.float32)
    x[0, 0] = [np.nan, 3.0, 1000.0]  # NaN slot, table slot, beyond-table slot
    x[0, 1] = [-4.0, 0.5, 300.0]
    mod = PositionEmbeddingSine3D(3584)
    got = mod(torch.from_numpy(x).to(DEV)).cpu().numpy()
    ref = O.pe_sine3d(x, 3584)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    np.testing.assert_allclose(np.nan_to_num(got), np.nan_to_num(ref), rtol=0, atol=2e-6)


def test_pe_module_fused_add():
    rng = np.random.default_rng(6)
    x = rng.integers(0, 300, (2, 196, 3)).astype(np.float32)
    tokens = synthetic.make_tokens(3, 2, 196, 3584)
    mod = PositionEmbeddingSine3D(3584)
    got = mod(torch.from_numpy(x).to(DEV).to(torch.bfloat16), add_to=tokens.to(DEV))
    xb = torch.from_numpy(x).to(torch.bfloat16).float().numpy()
    ref = O.add_pe_bf16(bf16_bits(tokens), O.pe_sine3d(xb, 3584))
    assert_bf16_close(bf16_bits(got), ref, max_frac=5e-4)


def test_pe_golden(G):
    mod = PositionEmbeddingSine3D(3584)
    got = mod(torch.from_numpy(G["disc27"].reshape(2, 196, 3)).to(DEV))
    np.testing.assert_allclose(got[:, ::13].cpu().numpy(), G["pe3584_sel"], rtol=0, atol=5e-7)
    got = PositionEmbeddingSine3D(1155)(torch.from_numpy(G["disc14"].reshape(2, 729, 3)).to(DEV))
    np.testing.assert_allclose(got[:, ::37].cpu().numpy(), G["pe1155_sel"], rtol=0, atol=5e-7)
    got = PositionEmbeddingSine3D(3584, n_points=9)(torch.from_numpy(G["disc_sample9"].reshape(2, 196, 9, 3)).to(DEV))
    np.testing.assert_allclose(got[:, ::13].cpu().numpy(), G["pe3584_n9_sel"], rtol=0, atol=5e-7)
    got = PositionEmbeddingSine3D(3584)(torch.from_numpy(G["avg27"].reshape(2, 196, 3)).to(DEV))
    np.testing.assert_allclose(got[:, ::13].cpu().numpy(), G["pe3584_cont_sel"], rtol=0, atol=2e-6)


# ---------------------------------------------------------------------------------------------
# the fused kernel
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("patch,E", [(27, 3584), (14, 1152)])
@pytest.mark.parametrize("add", [False, True])
def test_fused_frontend_default_recipe(scene, patch, E, add):
    d, p, k = _dev(scene)
    fe = FrontEnd(device=DEV, patch_size=patch, embedding_size=E)
    n_tok = fe.tokens_per_frame
    tokens = synthetic.make_tokens(11, 3, n_tok, E) if add else None
    out, pooled, pe_coords = fe(d, p, k, tokens=None if tokens is None else tokens.to(DEV), want_coords=True)
    ref = O.frontend(scene.depth_u16(), scene.poses.numpy(), scene.intrinsics.numpy(), patch_size=patch, embedding_size=E,
                     tokens_bf16_bits=None if tokens is None else bf16_bits(tokens))
    np.testing.assert_array_equal(pooled.cpu().numpy(), ref["pooled"])
    np.testing.assert_array_equal(pe_coords.cpu().numpy().reshape(3, n_tok, 3), ref["pe_coords"])
    assert_bf16_close(bf16_bits(out), ref["tokens_out"] if add else ref["pe_bf16"], max_frac=5e-4)


def test_fused_in_place_add(scene):
    d, p, k = _dev(scene)
    fe = FrontEnd(device=DEV, patch_size=14, embedding_size=1152)
    tokens = synthetic.make_tokens(12, 3, 729, 1152).to(DEV)
    expect = fe(d, p, k, tokens=tokens)
    buf = tokens.clone()
    got = fe(d, p, k, tokens=buf, out=buf)
    assert got.data_ptr() == buf.data_ptr()
    assert torch.equal(got.view(torch.int16), expect.view(torch.int16))


@pytest.mark.parametrize("kind,n_points", [("minmax-discrete-sin3d", 2), ("sample9-discrete-sin3d", 9), ("sample5-discrete-sin3d", 5),
                                           ("sample1-discrete-sin3d", 1), ("avg-sin3d", 1), ("minmax-sin3d", 2)])
def test_fused_frontend_variants(scene, kind, n_points):
    d, p, k = _dev(scene)
    fe = FrontEnd(device=DEV, patch_size=27, embedding_size=3584, world_position_embedding_type=kind)
    out, pooled, pe_coords = fe(d, p, k, want_coords=True)
    pool = kind.split("-")[0]
    ref = O.frontend(scene.depth_u16(), scene.poses.numpy(), scene.intrinsics.numpy(), patch_size=27, embedding_size=3584,
                     pool=pool, discrete="discrete" in kind)
    np.testing.assert_array_equal(pooled.cpu().numpy(), ref["pooled"])
    np.testing.assert_array_equal(pe_coords.cpu().numpy().reshape(ref["pe_coords"].shape), ref["pe_coords"])
    if "discrete" in kind:
        assert_bf16_close(bf16_bits(out), ref["pe_bf16"], max_frac=5e-4)
    else:
        got = out.float().cpu().numpy()
        np.testing.assert_allclose(got, O.bf16_bits_to_f32(ref["pe_bf16"]), rtol=0, atol=1.0 / 128)
        assert float(np.mean(bf16_bits(out) != ref["pe_bf16"])) < 2e-3


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_fused_frontend_other_dtypes(scene, dtype):
    d, p, k = _dev(scene)
    fe = FrontEnd(device=DEV, patch_size=27, embedding_size=3584, dtype=dtype)
    out = fe(d, p, k)
    ref = O.frontend(scene.depth_u16(), scene.poses.numpy(), scene.intrinsics.numpy(), patch_size=27, embedding_size=3584)
    if dtype == torch.float32:
        np.testing.assert_allclose(out.cpu().numpy(), ref["pe_f32"], rtol=0, atol=5e-7)
    else:
        np.testing.assert_allclose(out.float().cpu().numpy(), ref["pe_f32"], rtol=0, atol=1e-3)


def test_fused_frontend_norm_clamp_and_resize_strategy(scene):
    d, p, k = _dev(scene)
    lo, hi = [-1.0, -1.5, 0.2], [1.2, 1.0, 2.0]
    fe = FrontEnd(device=DEV, patch_size=27, embedding_size=3584, strategy="resize")
    out, pooled, pe_coords = fe(d, p, k, clamp=(lo, hi), want_coords=True)
    ref = O.frontend(scene.depth_u16(), scene.poses.numpy(), scene.intrinsics.numpy(), patch_size=27, embedding_size=3584,
                     strategy="resize", clamp_lo=lo, clamp_hi=hi)
    np.testing.assert_array_equal(pooled.cpu().numpy(), ref["pooled"])
    assert_bf16_close(bf16_bits(out), ref["pe_bf16"], max_frac=5e-4)


@pytest.mark.parametrize("depth_kind", ["i32", "f32", "u16_unaligned_width"])
def test_fused_frontend_depth_dtypes_and_non_tma_path(depth_kind):
    if depth_kind == "u16_unaligned_width":
        sc = synthetic.make_scene(4, 2, n_total=10, H=481, W=643, device="cpu")  # W*2 % 16 != 0 -> plain-load variant
    else:
        sc = synthetic.make_scene(4, 2, n_total=10, device="cpu")
    du = sc.depth_u16()
    d = sc.depth
    if depth_kind == "i32":
        d = torch.from_numpy(du.astype(np.int32))
    elif depth_kind == "f32":
        d = torch.from_numpy(du.astype(np.float32))
    H, W = du.shape[1:]
    fe = FrontEnd(device=DEV, H=H, W=W, patch_size=14, embedding_size=1152)
    out, pooled, _ = fe(d.to(DEV), sc.poses.to(DEV), sc.intrinsics.to(DEV), want_coords=True)
    ref = O.frontend(du, sc.poses.numpy(), sc.intrinsics.numpy(), patch_size=14, embedding_size=1152)
    np.testing.assert_array_equal(pooled.cpu().numpy(), ref["pooled"])
    assert_bf16_close(bf16_bits(out), ref["pe_bf16"], max_frac=5e-4)


def test_fused_frontend_golden(G):
    """Against outputs of the reference itself (tests/golden/frontend.npz)."""
    sc = synthetic.make_scene(0, 2, n_total=30, device="cpu")
    tokens = synthetic.make_tokens(0, 2, 196, 3584)
    fe = FrontEnd(device=DEV, patch_size=27, embedding_size=3584)
    out, pooled, pe_coords = fe(sc.depth.to(DEV), sc.poses.to(DEV), sc.intrinsics.to(DEV), tokens=tokens.to(DEV), want_coords=True)
    np.testing.assert_array_equal(pooled.cpu().numpy(), G["avg27"])
    np.testing.assert_array_equal(pe_coords.cpu().numpy(), G["disc27"])
    assert_bf16_close(bf16_bits(out[:, ::13]), O.f32_to_bf16_bits(G["added3584_sel"]), max_frac=5e-4)
    fe14 = FrontEnd(device=DEV, patch_size=14, embedding_size=1152)
    out14, pooled14, pc14 = fe14(sc.depth.to(DEV), sc.poses.to(DEV), sc.intrinsics.to(DEV), want_coords=True)
    np.testing.assert_array_equal(pooled14.cpu().numpy(), G["avg14"])
    np.testing.assert_array_equal(pc14.cpu().numpy(), G["disc14"])
    assert_bf16_close(bf16_bits(out14[:, ::37]), O.f32_to_bf16_bits(G["pe1152_sel"]), max_frac=5e-4)


def test_fused_full_size_properties():
    """BASELINE config 2 size (16 scenes x 32 frames, 27x27 grid, E=1152): properties that do not need
    the oracle at full size -- linearity of the add, determinism, agreement with the split path."""
    scenes = [synthetic.make_scene(100 + s, 32, n_total=300, device=DEV) for s in range(16)]
    depth = torch.cat([s.depth for s in scenes])
    poses = torch.cat([s.poses for s in scenes])
    intr = torch.cat([s.intrinsics for s in scenes])
    fe = FrontEnd(device=DEV, patch_size=14, embedding_size=1152)
    pe, pooled, pe_coords = fe(depth, poses, intr, want_coords=True)
    assert pe.shape == (512, 729, 1152)
    assert torch.equal(pe.view(torch.int16), fe(depth, poses, intr).view(torch.int16))  # deterministic
    zeros = torch.zeros_like(pe)
    assert torch.equal(fe(depth, poses, intr, tokens=zeros).view(torch.int16), pe.view(torch.int16))  # PE + 0 == PE
    tok = torch.randn(pe.shape, device=DEV, dtype=torch.bfloat16)
    fused = fe(depth, poses, intr, tokens=tok)
    assert torch.equal(fused.view(torch.int16), (tok + pe).view(torch.int16))  # fused add == torch's bf16 add
    # split path: the stand-alone PE kernel on the discretised coordinates gives the same rows
    pe3 = PositionEmbeddingSine3D(1152)(pe_coords.reshape(512, 729, 3))
    assert torch.equal(pe3.to(torch.bfloat16).view(torch.int16), pe.view(torch.int16))
    # first scene against the oracle
    s0 = scenes[0]
    ref = O.frontend(s0.depth_u16(), s0.poses.cpu().numpy(), s0.intrinsics.cpu().numpy(), patch_size=14, embedding_size=1152)
    np.testing.assert_array_equal(pooled[:32].cpu().numpy(), ref["pooled"])
    assert_bf16_close(bf16_bits(pe[:32]), ref["pe_bf16"], max_frac=5e-4)


def test_errors_are_reported_not_swallowed(scene):
    d, p, k = _dev(scene)
    with pytest.raises(RuntimeError):
        FrontEnd(device="cpu")
    with pytest.raises(RuntimeError):
        C.average_coordinate_in_patch(torch.zeros(1, 384, 384, 3))  # CPU tensor: no CPU path
    fe = FrontEnd(device=DEV, patch_size=27, embedding_size=3584)
    with pytest.raises(ValueError):
        fe(d[:, :100], p, k)
    with pytest.raises(ValueError):
        fe(d, p, k, tokens=torch.zeros(3, 196, 3584, device=DEV))  # wrong dtype
    with pytest.raises(NotImplementedError):
        PositionEmbeddingSine3D(12)(torch.zeros(1, 4, 3, device=DEV))  # nf < 8 unsupported, reported as such
    lib = cabi.load()
    assert lib.v3d_unproject(None, 0, None, None, 1, 1, 1, 1000.0, None, None, 0, 0, None, None, None) == cabi.V3D_E_ARG
    assert b"null pointer" in lib.v3d_last_error()


# ---------------------------------------------------------------------------------------------
# MC sampler
# ---------------------------------------------------------------------------------------------
def _mc_scene(seed, n_frames, H, W):
    sc = synthetic.make_scene(seed, n_frames, n_total=n_frames, H=H, W=W, device="cpu")
    pc = synthetic.mesh_voxels(sc.room, sc.boxes, seed)
    return sc, pc


def test_mc_voxel_ids(scene):
    world = O.unproject(scene.intrinsics.numpy(), scene.poses.numpy(), scene.depth_u16())[:1]
    world[0, 0, 0] = [np.nan, np.inf, -3e12]
    for mode_o, mode_c in ((O.DIV_IEEE, cabi.DIV_IEEE), (O.DIV_RECIPROCAL, cabi.DIV_RECIPROCAL)):
        got = mc_sampling.voxel_ids(torch.from_numpy(world).to(DEV), 0.1, mode_c).cpu().numpy()
        np.testing.assert_array_equal(got, O.mc_voxel_ids(world, 0.1, mode_o))


@pytest.mark.parametrize("seed,n_frames,H,W", [(21, 40, 120, 160), (22, 9, 480, 640), (23, 70, 60, 80)])
def test_mc_sampler_bit_exact_vs_oracle(seed, n_frames, H, W):
    sc, pc = _mc_scene(seed, n_frames, H, W)
    sampler = mc_sampling.MCSampler(device=DEV)
    plan = sampler.plan([{"depth": sc.depth.to(DEV), "poses": sc.poses.to(DEV), "intrinsics": sc.intrinsics.to(DEV),
                          "pc_voxels": torch.from_numpy(pc).to(DEV), "seed": seed}])
    res = sampler.decode(sampler.run(plan))[0]
    order, gains, n_all, n_sel, bs = O.mc_sample_scene(sc.depth_u16(), sc.poses.numpy(), sc.intrinsics.numpy(), pc, 0.1, 32, seed)
    np.testing.assert_array_equal(sampler.bitsets(plan, 0), bs)  # voxel coverage bit-exact
    assert res["order"] == list(order)  # including MT19937 tie-breaks
    assert res["gains"] == list(gains)
    assert (res["n_all"], res["n_sel"]) == (n_all, n_sel)


@pytest.mark.parametrize("case", ["half_voxel_shift", "odd_width", "int32_depth", "mixed_pose_kinds", "reciprocal", "huge_translation"])
def test_mc_voxelize_guard_band_and_generic_paths(case):
    """The voxeliser accepts an FMA estimate of the voxel id only outside a guard band around the half-integers
    and runs the reference operator sequence otherwise (and for every frame that is not uint16 / rigid / 8-pixel
    aligned).  Walls shifted onto half-voxel planes put 2-5 % of the pixels into the band."""
    H, W = (60, 84) if case == "odd_width" else (240, 320)
    sc = synthetic.make_scene(41, 6, n_total=6, H=H, W=W, device="cpu")
    poses = sc.poses.clone()
    depth = sc.depth
    div_o, div_c = O.DIV_IEEE, cabi.DIV_IEEE
    if case == "half_voxel_shift":
        poses[:, :3, 3] += torch.tensor([0.05, 0.05, 0.05])
    if case == "huge_translation":  # |t| beyond the range of the rounding trick: whole frames take the exact path
        poses[3:, :3, 3] += torch.tensor([3.0e5, 0.0, 0.0])
        poses[:3, :3, 3] += torch.tensor([1.0e4, 0.05, 0.0])
    if case == "mixed_pose_kinds":
        poses[1, 3, :] = torch.tensor([0.0, 0.0, 1e-3, 1.0])  # projective bottom row
        poses[4, 0, 0] = float("nan")
    if case == "int32_depth":
        depth = torch.from_numpy(sc.depth_u16().astype(np.int32))
    if case == "reciprocal":
        div_o, div_c = O.DIV_RECIPROCAL, cabi.DIV_RECIPROCAL
        poses[:, :3, 3] += torch.tensor([0.05, -0.05, 0.25])
    d_np = sc.depth_u16() if case != "int32_depth" else depth.numpy()
    world = O.unproject(sc.intrinsics.numpy(), poses.numpy(), d_np)
    ids = O.mc_voxel_ids(world, 0.1, div_o)
    uni = np.unique(ids[:, ::3, ::3].reshape(-1, 3), axis=0)
    rng = np.random.default_rng(7)
    pc = np.ascontiguousarray(uni[rng.permutation(len(uni))][: max(1, int(0.9 * len(uni)))]).astype(np.int32)
    sampler = mc_sampling.MCSampler(device=DEV, div_mode=div_c)
    plan = sampler.plan([{"depth": depth.to(DEV), "poses": poses.to(DEV), "intrinsics": sc.intrinsics.to(DEV),
                          "pc_voxels": torch.from_numpy(pc).to(DEV), "seed": 5}])
    sampler.run(plan, "voxelize")
    np.testing.assert_array_equal(sampler.bitsets(plan, 0), O.mc_bitsets(ids, pc))


def test_mc_sampler_batch_ragged_and_edge_cases():
    specs = [(31, 33, 120, 160), (32, 5, 120, 160), (33, 1, 120, 160), (34, 48, 120, 160)]
    scenes, refs = [], []
    for i, (seed, n, H, W) in enumerate(specs):
        sc, pc = _mc_scene(seed, n, H, W)
        if i == 1:
            pc = pc[:0]  # empty universe: every gain is 0, all picks are tie-breaks
        if i == 3:
            pc = np.concatenate([pc, pc[:50]])  # duplicated voxels behave like a set
        scenes.append({"depth": sc.depth.to(DEV), "poses": sc.poses.to(DEV), "intrinsics": sc.intrinsics.to(DEV),
                       "pc_voxels": torch.from_numpy(np.ascontiguousarray(pc)).to(DEV), "seed": 2**40 + seed})
        upc = pc if i != 3 else pc[: len(pc) - 50]
        refs.append(O.mc_sample_scene(sc.depth_u16(), sc.poses.numpy(), sc.intrinsics.numpy(), upc, 0.1, 32, 2**40 + seed))
    res = mc_sampling.MCSampler(device=DEV).sample(scenes)
    for r, (order, gains, n_all, n_sel, _) in zip(res, refs):
        assert r["order"] == list(order) and r["gains"] == list(gains) and (r["n_all"], r["n_sel"]) == (n_all, n_sel)


def test_mc_full_size_properties():
    """BASELINE config 3 size (300-frame scenes -> 150 candidates of 480x640, 0.1 m voxels): properties that do not need
    the oracle at full size.  The same scenes as int32 depth take the exact operator sequence for every pixel (no FMA
    estimate, no guard band): coverage bitsets and picks must be identical to the uint16 fast path."""
    scenes16, scenes32 = [], []
    for s in range(4):
        seed = 700 + s
        sc = synthetic.make_scene(seed, 150, n_total=300, frame_indices=np.arange(0, 300, 2), device=DEV)
        pc = torch.from_numpy(synthetic.mesh_voxels(sc.room, sc.boxes, seed)).to(DEV)
        base = {"poses": sc.poses, "intrinsics": sc.intrinsics, "pc_voxels": pc, "seed": seed}
        scenes16.append(dict(base, depth=sc.depth))
        scenes32.append(dict(base, depth=(sc.depth.to(torch.int32) & 0xFFFF)))
    sampler = mc_sampling.MCSampler(device=DEV)
    plan16, plan32 = sampler.plan(scenes16), sampler.plan(scenes32)
    res16 = sampler.decode(sampler.run(plan16))
    res32 = sampler.decode(sampler.run(plan32))
    again = sampler.decode(sampler.run(plan16))
    assert res16 == again  # deterministic, and re-running a plan clears its workspace
    for i, (a, b) in enumerate(zip(res16, res32)):
        np.testing.assert_array_equal(sampler.bitsets(plan16, i), sampler.bitsets(plan32, i))
        assert a == b
        bits = sampler.bitsets(plan16, i)
        union = np.bitwise_or.reduce(bits, axis=0)
        popc = lambda w: int(np.unpackbits(w.view(np.uint8)).sum())
        assert a["n_all"] == popc(union)  # |union of frames & pc| (max_coverage_sampling.py:102)
        assert len(a["order"]) == 32 and len(set(a["order"])) == 32 and all(0 <= f < 150 for f in a["order"])
        assert all(x >= y for x, y in zip(a["gains"], a["gains"][1:]))  # greedy gains of a coverage function never grow
        assert sum(a["gains"]) == a["n_sel"] <= a["n_all"]
        used = np.zeros_like(union)
        for f, g in zip(a["order"], a["gains"]):  # replay: every recorded gain is the marginal coverage of its pick
            assert g == popc(bits[f] & ~used)
            assert g == max(popc(bits[k] & ~used) for k in range(150) if k not in a["order"][: a["order"].index(f)])
            used |= bits[f]


def test_mc_full_size_scene_against_the_oracle():
    """One BASELINE config 3 scene at full size (150 candidate frames of 480x640, 0.1 m voxels) against the C oracle
    (oracle/v3d_oracle.c restates max_coverage_sampling.py:39-104 with every rounding explicit; pinned on the reference's own
    output in tests/test_oracle_golden.py): coverage bitsets of all 150 frames, pick order with the seeded tie-breaks, gains
    and both counts must be identical."""
    seed = 704
    sc = synthetic.make_scene(seed, 150, n_total=300, frame_indices=np.arange(0, 300, 2), device="cpu")
    pc = synthetic.mesh_voxels(sc.room, sc.boxes, seed)
    order, gains, n_all, n_sel, bitsets = O.mc_sample_scene(sc.depth_u16(), sc.poses.numpy(), sc.intrinsics.numpy(), pc, 0.1, 32, seed)
    sampler = mc_sampling.MCSampler(device=DEV)
    plan = sampler.plan([{"depth": sc.depth.to(DEV), "poses": sc.poses.to(DEV), "intrinsics": sc.intrinsics.to(DEV),
                          "pc_voxels": torch.from_numpy(pc).to(DEV), "seed": seed}])
    res = sampler.decode(sampler.run(plan))[0]
    got_bits = sampler.bitsets(plan, 0)
    ref_bits = np.asarray(bitsets).reshape(150, -1)
    np.testing.assert_array_equal(got_bits[:, : ref_bits.shape[1]].astype(np.uint32), ref_bits.astype(np.uint32))
    assert res["order"] == list(order) and res["gains"] == list(gains)
    assert (res["n_all"], res["n_sel"]) == (n_all, n_sel)
    assert len(order) == 32


def test_mc_sampler_golden(golden_dir):
    """Against the output of the reference's own max_coverage_sampling.main (tests/golden/dataset.json)."""
    with open(os.path.join(golden_dir, "dataset.json")) as f:
        D = json.load(f)
    for name, meta in D["scenes"].items():
        sc = synthetic.make_scene(meta["seed"], meta["n_total"], n_total=meta["n_total"], H=meta["H"], W=meta["W"], device="cpu")
        pc = synthetic.mesh_voxels(sc.room, sc.boxes, meta["seed"])
        cand = fixtures.mc_candidates(meta["n_total"])
        poses = torch.from_numpy(fixtures.disk_poses(sc))[cand]
        gold = next(s for s in D["select_frames"] if s["video_id"].endswith(name))
        res = mc_sampling.MCSampler(device=DEV).sample([{"depth": sc.depth[cand].to(DEV), "poses": poses.to(DEV),
                                                         "intrinsics": sc.intrinsics[cand].to(DEV),
                                                         "pc_voxels": torch.from_numpy(pc).to(DEV), "seed": meta["seed"]}])[0]
        assert [cand[i] for i in res["order"]] == [int(f.split("/")[-1].split(".")[0]) // 10 for f in gold["frame_files"]]
        assert res["gains"] == gold["voxel_nums"]
        assert (res["n_all"], res["n_sel"]) == (gold["num_all_voxels"], gold["num_select_voxels"])


# ---------------------------------------------------------------------------------------------
# the hoisted-reciprocal division is IEEE division, for every dividend
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("divisor", [1000.0, 577.87, 577.87 * 0.25, 0.1, 0.05, 518.8579, 1170.1877, -3.0, 1e-30, 3e38, 2.0 ** -45])
def test_exact_division_exhaustive(divisor):
    """div_exact(a, b) == a / b (div.rn.f32) for all 2^32 bit patterns of a: the depth scale, ScanNet-like
    focal lengths, voxel sizes, a negative divisor, and divisors outside the fast range (which must take
    the IEEE fallback)."""
    lib = cabi.load()
    bad = torch.zeros(1, dtype=torch.int64, device=DEV)
    cabi.check(lib.v3d_selftest_division(float(np.float32(divisor)), bad.data_ptr(), cabi.stream_ptr(DEV)))
    torch.cuda.synchronize()
    assert int(bad.item()) == 0


# ---------------------------------------------------------------------------------------------
# VideoProcessor end to end on the on-disk fixture, against the reference's own outputs
# ---------------------------------------------------------------------------------------------
def _vp(root, strategy, **kw):
    return VU.VideoProcessor(video_folder=os.path.join(root, "data"), annotation_dir=os.path.join(root, "data", "embodiedscan"),
                             voxel_size=0.1, min_xyz_range=[-15, -15, -5], max_xyz_range=[15, 15, 5],
                             frame_sampling_strategy=strategy, metadata_dir=os.path.join(root, "data", "metadata"), **kw)


@pytest.mark.parametrize("strategy", ["uniform", "mc-ratio95", "mc-ratio95-norm"])
@pytest.mark.parametrize("lazy", [False, True])
def test_video_processor_matches_reference(dataset, D, strategy, lazy, monkeypatch):
    import hashlib

    monkeypatch.chdir(dataset)  # the MC frame lists hold paths relative to the working directory, like upstream

    name = "scene0003_00"
    gold = D["video_processor"][strategy][name]
    vp = _vp(dataset, strategy, lazy=lazy)
    d = vp.process_3d_video(f"scannet/{name}", fixtures.FakeImageProcessor(384), force_sample=True, frames_upbound=3)
    wc = d["world_coords"]
    full = wc.materialize() if lazy else wc
    assert full.is_cuda and list(full.shape) == gold["world_coords_shape"]
    got = full.cpu().numpy()
    np.testing.assert_array_equal(got[:, ::61, ::67], np.array(gold["world_coords_sample"], np.float32))
    assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest() == gold["world_coords_sha"]
    np.testing.assert_array_equal(np.asarray(d["boundry"].tolist(), np.float32), np.array(gold["boundry"], np.float32))
    assert d["objects"].tolist() == gold["objects"] and d["video_size"] == gold["video_size"]


def test_online_mc_sampling_equals_the_offline_json(dataset, D, monkeypatch, tmp_path):
    """frame_sampling_strategy 'mc-online-ratio95' computes the max-coverage order inside
    VideoProcessor (device voxelise + greedy) and must give what the reference's offline script wrote to
    scannet_select_frames.json under the same tie-break seeds -- and the same sampled frames as 'mc-ratio95'.
    Also through the packed scene binaries (f3)."""
    from video_3d_llm_b200 import scene_pack

    monkeypatch.chdir(dataset)
    seeds = {f"scannet/{name}": m["seed"] for name, m in D["scenes"].items()}
    vp_off = _vp(dataset, "mc-ratio95")
    vp_on = _vp(dataset, "mc-online-ratio95", mc_seeds=seeds)
    packs = str(tmp_path)
    for video_id in vp_off.scene:
        scene_pack.pack_scene(vp_off, video_id, scene_pack.pack_path(packs, video_id))
    vp_on_pack = _vp(dataset, "mc-online-ratio95", mc_seeds=seeds, pack_dir=packs)
    for gold in D["select_frames"]:
        vid = gold["video_id"]
        for vp in (vp_on, vp_on_pack):
            rec = vp.mc_sample_online(vid)
            # the offline JSON holds paths relative to the working directory (video_folder "data"), as upstream
            assert [os.path.abspath(f) for f in rec["frame_files"]] == [os.path.abspath(f) for f in gold["frame_files"]]
            assert rec["voxel_nums"] == gold["voxel_nums"]
            assert (rec["num_all_voxels"], rec["num_select_voxels"]) == (gold["num_all_voxels"], gold["num_select_voxels"])
            for ub in (3, 8, 32):
                assert [os.path.abspath(f) for f in vp.sample_frame_files_mc(vid, frames_upbound=ub)] == \
                       [os.path.abspath(f) for f in vp_off.sample_frame_files_mc(vid, frames_upbound=ub)]
    ip = fixtures.FakeImageProcessor(384)
    a = vp_off.process_3d_video("scannet/scene0003_00", ip, force_sample=True, frames_upbound=3)
    b = vp_on_pack.process_3d_video("scannet/scene0003_00", ip, force_sample=True, frames_upbound=3)
    assert torch.equal(a["world_coords"].materialize(), b["world_coords"].materialize())
    assert torch.equal(a["images"], b["images"])


def test_lazy_handle_through_the_reference_call_structure(dataset, D, monkeypatch):
    monkeypatch.chdir(dataset)
    """What prepare_inputs_labels_for_multimodal does with video_dict (llava_arch.py:386-420, 506-517), on a
    lazy handle: per-scene pooling straight from depth, discretisation, PE + add -- equal to the eager path."""
    vp_l, vp_e = _vp(dataset, "uniform", lazy=True), _vp(dataset, "uniform", lazy=False)
    ip = fixtures.FakeImageProcessor(384)
    dl = vp_l.process_3d_video("scannet/scene0003_00", ip, force_sample=True, frames_upbound=3)
    de = vp_e.process_3d_video("scannet/scene0003_00", ip, force_sample=True, frames_upbound=3)
    dl["box_input"] = de["box_input"] = None
    ml, me = VU.merge_video_dict([dl]), VU.merge_video_dict([de])
    wl = ml["world_coords"].to(DEV).to(torch.bfloat16)  # Trainer moves + casts float tensors to bf16
    we = me["world_coords"].to(DEV)
    _ = wl.view(1, -1, 3).min(dim=1)[0]  # llava_arch.py:387: accepted, never evaluated
    pooled_l = [C.average_coordinate_in_patch(c) for c in wl]
    # the reference pools the bf16 tensor: every coordinate is rounded to bf16 BEFORE the window sums
    pooled_e = [C.average_coordinate_in_patch(c) for c in we.to(torch.bfloat16)]
    assert pooled_l[0].dtype == torch.bfloat16 and pooled_l[0].shape == (3, 14, 14, 3)
    assert torch.equal(pooled_l[0].float(), pooled_e[0].float())
    disc = C.discrete_coords(pooled_e[0], None)
    mod = PositionEmbeddingSine3D(3584)
    tokens = synthetic.make_tokens(5, 3, 196, 3584).to(DEV)
    out = tokens + mod(disc.flatten(1, 2).to(torch.bfloat16))
    assert out.shape == (3, 196, 3584)
    # slicing like the object-proposal code (llava_arch.py:360-373) materialises on demand
    sl = ml["world_coords"].to(DEV)[0][:, :378, :378, :]
    assert torch.equal(sl, we[0][:, :378, :378, :])


# ---------------------------------------------------------------------------------------------
# the offline scripts: mesh universe and the MC command line (same files in, same JSON out)
# ---------------------------------------------------------------------------------------------
def test_mesh_to_voxels_matches_convert_pcd_to_voxel():
    """convert_pcd_to_voxel.py:69-82: unique round(xyz / 0.1) of the mesh vertices (CPU IEEE division)."""
    rng = np.random.default_rng(3)
    xyz = rng.uniform(-6, 6, (200000, 3)).astype(np.float32)
    xyz[:5] = [[0.05, 0.15, -0.25], [0.0, 0.0, 0.0], [-0.05, 2.25, 1.35], [5.95, -5.95, 0.049999], [1e-9, -1e-9, 3.0]]
    got = sorted(mc_sampling.mesh_to_voxels(torch.from_numpy(xyz).to(DEV), 0.1))
    ref = sorted(set(map(tuple, np.rint(xyz / np.float32(0.1)).astype(np.int32).tolist())))
    assert got == ref


def test_mc_command_line_reproduces_the_reference_json(dataset, D, monkeypatch, tmp_path):
    """`python -m video_3d_llm_b200.mc_sampling` on the on-disk fixture writes what the reference's
    max_coverage_sampling.main wrote (tests/golden/dataset.json), scene by scene under random.seed(scene seed)."""
    import types

    monkeypatch.chdir(dataset)
    gold = {s["video_id"]: s for s in D["select_frames"]}
    for name, meta in D["scenes"].items():
        vd = tmp_path / f"only_{name}"
        (vd / name).mkdir(parents=True)
        args = types.SimpleNamespace(voxel_size=0.1, video_dir=str(vd), output_file=str(tmp_path / f"sel_{name}.json"),
                                     seed=meta["seed"], batch_scenes=4, div_mode="ieee")  # the golden JSON is the reference on the CPU
        out = mc_sampling.main(args)
        with open(args.output_file) as f:
            assert json.load(f) == out
        assert len(out) == 1
        g = gold[f"scannet/{name}"]
        assert out[0]["video_id"] == g["video_id"]
        assert out[0]["frame_files"] == g["frame_files"]
        assert out[0]["voxel_nums"] == g["voxel_nums"]
        assert out[0]["num_all_voxels"] == g["num_all_voxels"] and out[0]["num_select_voxels"] == g["num_select_voxels"]


# ---------------------------------------------------------------------------------------------
# empty / degenerate inputs
# ---------------------------------------------------------------------------------------------
def test_empty_and_degenerate_inputs():
    z = torch.zeros((0, 480, 640), dtype=torch.int16, device=DEV)
    m = torch.zeros((0, 4, 4), device=DEV)
    assert VU.unproject(m, m, z).shape == (0, 480, 640, 3)
    fe = FrontEnd(device=DEV, patch_size=14, embedding_size=1152)
    assert fe(z, m, m).shape == (0, 729, 1152)
    assert PositionEmbeddingSine3D(96)(torch.zeros((0, 5, 3), device=DEV)).shape == (0, 5, 96)
    assert C.discrete_coords(torch.zeros((0, 3), device=DEV)).shape == (0, 3)
    # a frame of all-invalid depth: every pixel unprojects to the camera centre, PE rows are all equal
    sc = synthetic.make_scene(2, 1, n_total=5, device="cpu")
    d0 = torch.zeros_like(sc.depth).to(DEV)
    out, pooled, pe_coords = fe(d0, sc.poses.to(DEV), sc.intrinsics.to(DEV), want_coords=True)
    ref = O.frontend(np.zeros((1, 480, 640), np.uint16), sc.poses.numpy(), sc.intrinsics.numpy(), patch_size=14, embedding_size=1152)
    np.testing.assert_array_equal(pooled.cpu().numpy(), ref["pooled"])
    assert torch.equal(out[0, 0], out[0, -1])
    # saturated depth (65535 mm) and a frame far outside the discretisation range clamp to the table's last row
    d1 = torch.full_like(sc.depth, -1).to(DEV)  # 0xFFFF
    out1, pooled1, pc1 = fe(d1, sc.poses.to(DEV), sc.intrinsics.to(DEV), want_coords=True)
    ref1 = O.frontend(np.full((1, 480, 640), 65535, np.uint16), sc.poses.numpy(), sc.intrinsics.numpy(), patch_size=14, embedding_size=1152)
    np.testing.assert_array_equal(pooled1.cpu().numpy(), ref1["pooled"])
    np.testing.assert_array_equal(pc1.cpu().numpy().reshape(1, 729, 3), ref1["pe_coords"])
    assert_bf16_close(bf16_bits(out1), ref1["pe_bf16"], max_frac=5e-4)


def test_nan_pose_propagates_like_the_reference():
    sc = synthetic.make_scene(6, 2, n_total=8, device="cpu")
    p = sc.poses.clone()
    p[1, 1, 3] = float("nan")
    fe = FrontEnd(device=DEV, patch_size=27, embedding_size=3584)
    out, pooled, pe_coords = fe(sc.depth.to(DEV), p.to(DEV), sc.intrinsics.to(DEV), want_coords=True)
    ref = O.frontend(sc.depth_u16(), p.numpy(), sc.intrinsics.numpy(), patch_size=27, embedding_size=3584)
    np.testing.assert_array_equal(pooled.cpu().numpy(), ref["pooled"])  # NaN in y of frame 1, everything else intact
    np.testing.assert_array_equal(pe_coords.cpu().numpy().reshape(2, 196, 3), ref["pe_coords"])
    got = out.float().cpu().numpy()
    exp = O.bf16_bits_to_f32(ref["pe_bf16"])
    assert np.array_equal(np.isnan(got), np.isnan(exp))
    assert np.isnan(got[1, :, 1194:2388]).all() and not np.isnan(got[0]).any()


def geometry_maps(H, W):
    from video_3d_llm_b200 import geometry

    rows, cols = geometry.resample_maps_host(H, W, 384, "center_crop")
    return list(rows), list(cols)


def test_windowed_depth_upload_feeds_the_same_result(scene):
    """FrontEnd.upload copies only the depth samples the kernel reads; the result equals a full upload."""
    fe = FrontEnd(device=DEV, patch_size=14, embedding_size=1152)
    assert fe.upload_window(2) == (160, 944, 5, 4)  # ScanNet geometry: columns 80..551, rows 4 of 5
    assert fe.upload_bytes(3) == 3 * 96 * 4 * 944
    h = scene.depth.pin_memory()
    poison = torch.full(h.shape, 12345, dtype=h.dtype, device=DEV)
    d = fe.upload(h, out=poison)
    torch.cuda.synchronize()
    got = d.cpu().view(3, 96, 5, 640)
    ref = scene.depth.view(3, 96, 5, 640)
    assert torch.equal(got[:, :, :4, 80:552], ref[:, :, :4, 80:552])
    assert bool((got[:, :, 4, :] == 12345).all()) and bool((got[:, :, :, :80] == 12345).all())  # untouched outside the window
    p, k = scene.poses.to(DEV), scene.intrinsics.to(DEV)
    a = fe(d, p, k)
    b = fe(scene.depth.to(DEV), p, k)
    assert torch.equal(a.view(torch.int16), b.view(torch.int16))
    # packed form: one contiguous H2D copy + device-side expansion gives the same frames inside the window
    assert fe.packed_shape(3) == (3, 384, 472)
    packed = fe.pack_window(scene.depth)
    assert packed.is_pinned() and torch.equal(packed.view(3, 96, 4, 472), ref[:, :, :4, 80:552])
    poison2 = torch.full(h.shape, 23456, dtype=h.dtype, device=DEV)
    d2 = fe.upload_packed(packed, out=poison2)
    torch.cuda.synchronize()
    got2 = d2.cpu().view(3, 96, 5, 640)
    assert torch.equal(got2[:, :, :4, 80:552], ref[:, :, :4, 80:552])
    assert bool((got2[:, :, 4, :] == 23456).all()) and bool((got2[:, :, :, 552:] == 23456).all())
    assert torch.equal(fe(d2, p, k).view(torch.int16), b.view(torch.int16))
    with pytest.raises(ValueError):
        fe.upload_packed(packed[:, :100].contiguous().pin_memory())
    # resampled form: the loader applies the nearest resize + crop to the depth; scattered back on the device
    res = fe.pack_resampled(scene.depth)
    assert fe.resampled_shape(3) == (3, 384, 384) and tuple(res.shape) == (3, 384, 384) and res.is_pinned()
    rows_h, cols_h = geometry_maps(480, 640)
    assert torch.equal(res, scene.depth[:, rows_h][:, :, cols_h])
    poison3 = torch.full(h.shape, 3456, dtype=h.dtype, device=DEV)
    d3 = fe.scatter_resampled(res.to(DEV), poison3)
    torch.cuda.synchronize()
    touched = torch.zeros(h.shape, dtype=torch.bool)
    touched[:, torch.tensor(rows_h)[:, None], torch.tensor(cols_h)[None, :]] = True
    assert torch.equal(d3.cpu()[touched], scene.depth[touched]) and bool((d3.cpu()[~touched] == 3456).all())
    out_r, pooled_r, _ = fe(d3, p, k, want_coords=True)
    out_f, pooled_f, _ = fe(scene.depth.to(DEV), p, k, want_coords=True)
    assert torch.equal(out_r.view(torch.int16), out_f.view(torch.int16)) and torch.equal(pooled_r, pooled_f)
    # the fused kernel reads the resampled form in place (v3d_frontend_desc.depth_resampled): 768-byte rows staged by TMA ...
    out_i, pooled_i, disc_i = fe.from_resampled(res, scene.poses, scene.intrinsics, want_coords=True)
    assert torch.equal(out_i.view(torch.int16), out_f.view(torch.int16)) and torch.equal(pooled_i, pooled_f)
    # ... a 378 x 378 stack (756-byte rows: plain loads), both geometries, and the sample modes
    res378 = res[:, :378, :378].contiguous().to(DEV)
    assert torch.equal(fe(res378, p, k, resampled=True).view(torch.int16), out_f.view(torch.int16))
    for kw in (dict(patch_size=14, embedding_size=1152), dict(patch_size=27, embedding_size=1152, world_position_embedding_type="sample9-discrete-sin3d"),
               dict(patch_size=27, embedding_size=1152, world_position_embedding_type="minmax-sin3d")):
        fe_k = FrontEnd(device=DEV, **kw)
        full = fe_k(scene.depth.to(DEV), p, k, want_coords=True, clamp=((-2, -2, 0), (2, 2, 2)))
        for stack in (res.to(DEV), res378):
            got = fe_k(stack, p, k, want_coords=True, clamp=((-2, -2, 0), (2, 2, 2)), resampled=True)
            for x, y in zip(got, full):
                assert torch.equal(x.view(torch.int16) if x.dtype == torch.bfloat16 else x.view(torch.int32), y.view(torch.int16) if y.dtype == torch.bfloat16 else y.view(torch.int32))
    with pytest.raises(ValueError):
        fe(res[:, :300].contiguous().to(DEV), p, k, resampled=True)
    # the 'resize' strategy squeezes all 640 columns into 384: the window starts at column 0 (rows still 4 of 5)
    fe2 = FrontEnd(device=DEV, patch_size=27, embedding_size=3584, strategy="resize")
    x_off, width, period, keep = fe2.upload_window(2)
    assert (x_off, period, keep) == (0, 5, 4) and width <= 1280
    assert torch.equal(fe2(fe2.upload(h), p, k).view(torch.int16), fe2(scene.depth.to(DEV), p, k).view(torch.int16))


# ---------------------------------------------------------------------------------------------
# object proposals
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["patch14", "patch27"])
def test_object_patch_masks_bit_exact(golden_dir, kind):
    GO = np.load(os.path.join(golden_dir, "objects.npz"))
    sc = synthetic.make_scene(0, 2, n_total=30, device="cpu")
    from video_3d_llm_b200 import geometry
    rm, cm = geometry.resample_maps(480, 640, 384, "center_crop", DEV)
    coords = VU.unproject(sc.intrinsics, sc.poses, sc.depth, row_map=rm, col_map=cm)
    boxes = torch.from_numpy(GO["boxes"])
    got = C.object_patch_masks(coords, boxes, kind + "-pe").cpu().numpy()
    shape = tuple(GO[f"masks_{kind}_shape"])
    gold = np.unpackbits(GO[f"masks_{kind}"])[: int(np.prod(shape))].reshape(shape).astype(bool)
    np.testing.assert_array_equal(got, gold)  # the reference's own masks
    np.testing.assert_array_equal(got, O.object_patch_masks(coords.cpu().numpy(), GO["boxes"], kind))
    # a lazy handle gives the same masks straight from depth: the (V,384,384,3) tensor is never built
    lazy = VU.LazyWorldCoords(sc.depth, sc.poses, sc.intrinsics)
    assert torch.equal(C.object_patch_masks(lazy, boxes, kind), torch.from_numpy(gold).to(DEV))
    assert lazy._full is None
    # in the dtype the trainer / the eval scripts cast world_coords to, and with the 'norm' clamp: equal to the
    # materialised tensor of that dtype (the comparison runs on the rounded coordinates, as in the reference)
    for dt in (torch.bfloat16, torch.float16):
        for clamp in (None, ([-2.0, -1.5, 0.1], [1.5, 2.0, 1.9])):
            hz = VU.LazyWorldCoords(sc.depth.to(DEV), sc.poses.to(DEV), sc.intrinsics.to(DEV), clamp=clamp).to(dt)
            from_depth = C.object_patch_masks(hz, boxes, kind)
            assert hz._full is None
            assert torch.equal(from_depth, C.object_patch_masks(hz.materialize(), boxes, kind))


def test_object_features_match_reference(golden_dir):
    GO = np.load(os.path.join(golden_dir, "objects.npz"))
    shape = tuple(GO["masks_patch27_shape"])
    masks = torch.from_numpy(np.unpackbits(GO["masks_patch27"])[: int(np.prod(shape))].reshape(shape).astype(bool)).to(DEV)
    feats = synthetic.make_tokens(0, 2, 196, 256).to(DEV)
    got = C.object_features(feats, masks)
    assert got.dtype == torch.bfloat16 and got.shape == (shape[0], 256)
    np.testing.assert_allclose(got.float().cpu().numpy(), GO["objfeat_patch27"], rtol=0, atol=2.0 ** -7)
    empty = GO["masks_patch27_count"] == 0
    assert empty.any() and bool((got[torch.from_numpy(empty).to(DEV)] == 0).all())
    # larger, dense case against the oracle
    rng = np.random.default_rng(4)
    m = torch.from_numpy(rng.random((37, 2, 27, 27)) < 0.2).to(DEV)
    tok = synthetic.make_tokens(9, 2, 729, 1152).to(DEV)
    got2 = C.object_features(tok, m)
    ref2 = O.object_features(bf16_bits(tok).reshape(-1, 1152), m.cpu().numpy())
    assert_bf16_close(bf16_bits(got2), ref2, max_frac=5e-3)


@pytest.mark.parametrize("patch,E", [(27, 3584), (14, 1152), (27, 1152)])
def test_algebraic_patch_means_equal_the_operator_sequence(patch, E):
    """The streaming kernel (k_fused_ring) forms patch means from exact integer sums and verifies them against a proven
    error band (v3d_fused_ring.cu); coordinates inside the band, and frames whose camera does not qualify, are recomputed
    with the reference's operator sequence in the same launch.  The result must be what the operator-sequence kernel
    (k_fused, itself pinned on the oracle and the goldens) gives, bit for bit -- on ordinary scenes, far from the origin
    (large |P_i3|: wide band), on walls that sit exactly on voxel boundaries (every coordinate inside the band), with
    projective / NaN / huge cameras (nothing qualifies), with depth the loader already resampled, and in place."""
    fe = FrontEnd(device=DEV, patch_size=patch, embedding_size=E)
    fe_exact = FrontEnd(device=DEV, patch_size=patch, embedding_size=E)
    fe_exact.exact_pooling = True
    cases = 0
    for seed, kind in ((31, None), (32, "far"), (33, "wall"), (34, "projective"), (35, "nan_pose"), (36, "huge_focal"), (37, "resampled")):
        n_frames = 40 if kind is None else 7  # more patch rows than CTAs for one case, fewer for the others
        sc = synthetic.make_scene(seed, n_frames, n_total=40, device=DEV)
        poses, intr, depth = sc.poses.clone(), sc.intrinsics.clone(), sc.depth.clone()
        if kind == "far":
            poses[:, :3, 3] += torch.tensor((250.0, -180.0, 3.0), device=DEV)
        elif kind == "wall":
            # a fronto-parallel wall whose world z lands exactly on the half-voxel plane: camera looks down the world z axis
            poses[:] = torch.eye(4, device=DEV)
            poses[:, 2, 3] = 0.05
            depth[:] = 2000
        elif kind == "projective":
            poses[1::2, 3, :] = torch.tensor((0.01, -0.02, 0.03, 1.5), device=DEV)
        elif kind == "nan_pose":
            poses[2, 1, 2] = float("nan")
            poses[4, 0, 3] = float("inf")
        elif kind == "huge_focal":
            intr[3, 0, 0] = 3e30
            intr[5, 1, 1] = 1e-30
        tok = synthetic.make_tokens(seed, n_frames, fe.tokens_per_frame, E).to(DEV)
        kw = {}
        if kind == "resampled":
            depth = fe.pack_resampled(depth.cpu(), pin=False).to(DEV)
            kw = dict(resampled=True)
        a = fe(depth, poses, intr, tokens=tok, **kw)
        b = fe_exact(depth, poses, intr, tokens=tok, **kw)
        assert torch.equal(a.view(torch.int16), b.view(torch.int16)), (seed, kind)
        if kind is None:
            inplace = tok.clone()
            fe(depth, poses, intr, tokens=inplace, out=inplace)
            assert torch.equal(inplace.view(torch.int16), b.view(torch.int16))
        cases += 1
    assert cases == 7


def test_scene_batch_loader_pinned_cache_equals_the_host_path(dataset, tmp_path):
    """SceneBatchLoader(pinned_cache=True): scenes held in pinned memory, batches gathered by the DMA engines into device
    buffers (v3d_h2d_gather) -- the same bytes as the host-gather path, over more batches than ring slots (slot reuse waits
    for mark_consumed), and the fused front end gives the same tokens from either handle."""
    from video_3d_llm_b200 import scene_pack
    from video_3d_llm_b200 import video_utils as VU

    kw = dict(video_folder=os.path.join(dataset, "data"), annotation_dir=os.path.join(dataset, "data", "embodiedscan"), voxel_size=0.1,
              min_xyz_range=[-15, -15, -5], max_xyz_range=[15, 15, 5], frame_sampling_strategy="uniform",
              metadata_dir=os.path.join(dataset, "data", "metadata"))
    vp = VU.VideoProcessor(**kw)
    video_id = "scannet/scene0003_00"  # 480 x 640
    scene_pack.pack_scene(vp, video_id, scene_pack.pack_path(str(tmp_path), video_id), resample=(384, "center_crop"))
    vp_pack = VU.VideoProcessor(pack_dir=str(tmp_path), lazy=True, **kw)
    picks = [vp_pack.sample_frame_files(video_id, force_sample=True, frames_upbound=n) for n in (3, 3, 3)]
    all_files = vp_pack.sample_frame_files(video_id, force_sample=True, frames_upbound=4)
    samples = [(video_id, all_files[i:i + 3] if len(all_files[i:i + 3]) == 3 else all_files[:3]) for i in range(2)] + [(video_id, picks[0])]
    host = VU.SceneBatchLoader(vp_pack, samples, batch_scenes=1, threads=2, prefetch=2, loop=True, pin=True)
    dev = VU.SceneBatchLoader(vp_pack, samples, batch_scenes=1, threads=2, prefetch=2, loop=True, pinned_cache=True, device=DEV)
    fe = FrontEnd(device=DEV, patch_size=14, embedding_size=1152)
    tok = synthetic.make_tokens(5, 3, fe.tokens_per_frame, 1152).to(DEV)
    ih, idv = iter(host), iter(dev)
    for n in range(9):  # more than the 4 device ring slots
        a, b = next(ih), next(idv)
        assert b.depth.is_cuda and "_ready_event" in b.__dict__
        bd = b.to(DEV)  # waits for the loader's copy stream
        assert torch.equal(a.depth.to(DEV), bd.depth) and torch.equal(a.poses.to(DEV), bd.poses) and torch.equal(a.intrinsics.to(DEV), bd.intrinsics)
        out_a = fe.from_resampled(a.depth[0], a.poses[0], a.intrinsics[0], tokens=tok)
        out_b = fe.from_resampled(bd.depth[0], bd.poses[0], bd.intrinsics[0], tokens=tok)
        assert torch.equal(out_a.view(torch.int16), out_b.view(torch.int16))
        dev.mark_consumed(bd)
    host.close()
    dev.close()


def test_deferred_object_chain_falls_back_to_real_tensors():
    """lazy_objects.py records exactly the expressions of llava_arch.py:360-376; anything else on the same objects must behave
    like the materialised tensor (same values), and the recorded chain must equal the reference's expressions evaluated on it."""
    from video_3d_llm_b200 import video_utils as VU
    from video_3d_llm_b200.lazy_objects import BoxTest, ObjectMask, PatchGrid, PatchPixels

    sc = synthetic.make_scene(41, 3, n_total=30, device=DEV)
    h = VU.LazyWorldCoords(sc.depth, sc.poses, sc.intrinsics, src_hw=(480, 640))
    full = VU.LazyWorldCoords(sc.depth, sc.poses, sc.intrinsics, src_hw=(480, 640)).materialize()
    boxes = torch.tensor([[0.5, 0.2, 1.0, 2.5, 2.0, 2.0], [-1.0, 1.0, 0.8, 1.5, 1.5, 1.6], [9.0, 9.0, 9.0, 0.1, 0.1, 0.1]], device=DEV)
    ref_masks, got = [], []
    for patch, grid, thr in ((14, 27, int(14 * 14 * 0.5)), (27, 14, int(27 * 27 * 0.25))):
        for box in boxes:
            lo, hi = box[:3] - box[3:] / 2, box[:3] + box[3:] / 2
            w_ref = full[:, :378, :378, :].reshape(-1, grid, patch, grid, patch, 3).transpose(2, 3).flatten(3, 4)
            ref_masks.append(torch.all((lo <= w_ref) & (w_ref <= hi), dim=-1).sum(dim=3) >= thr)
            w = h[:, :378, :378, :]
            assert isinstance(w, PatchGrid)
            w = w.reshape(-1, grid, patch, grid, patch, 3).transpose(2, 3).flatten(3, 4)
            assert isinstance(w, PatchPixels)
            t = (lo <= w) & (w <= hi)
            assert isinstance(t, BoxTest)
            m = torch.all(t, dim=-1).sum(dim=3) >= thr
            assert isinstance(m, ObjectMask) and m._val is None
            got.append(m)
    assert h._full is None  # nothing materialised so far
    for m, r in zip(got, ref_masks):
        assert torch.equal(m.view(-1, r.shape[-1] * r.shape[-2]), r.view(-1, r.shape[-1] * r.shape[-2]))
    assert h._full is None and any(bool(r.any()) for r in ref_masks)
    # unmodelled uses: same values as the tensor expressions
    w = h[:, :378, :378, :]
    assert torch.equal(w.reshape(-1, 3).sum(0), full[:, :378, :378, :].reshape(-1, 3).sum(0))
    wp = h[:, :378, :378, :].reshape(-1, 27, 14, 27, 14, 3).transpose(2, 3).flatten(3, 4)
    lo = torch.tensor([0.0, 0.0, 0.0], device=DEV)
    t_ref = full[:, :378, :378, :].reshape(-1, 27, 14, 27, 14, 3).transpose(2, 3).flatten(3, 4)
    assert torch.equal((lo <= wp).float().mean(), (lo <= t_ref).float().mean())           # a lone comparison, then tensor methods
    assert torch.equal(torch.all(lo <= wp, dim=-1).sum(), torch.all(lo <= t_ref, dim=-1).sum())  # all() of a half-open test
    assert torch.equal((wp <= 1.5).sum(), (t_ref <= 1.5).sum())                                # scalar bound: not a box
    assert torch.equal(h[:, :100, :378, :], full[:, :100, :378, :])                             # another slice: plain indexing


@pytest.mark.parametrize("patch,E,dtype", [(42, 1152, torch.bfloat16), (54, 1152, torch.bfloat16), (63, 96, torch.float16), (14, 8192, torch.bfloat16),
                                           (27, 64, torch.bfloat16), (18, 3584, torch.float16)])
def test_streaming_kernel_geometries_and_fallback(patch, E, dtype):
    """Geometries around the edges of what k_fused_ring takes (few ring slots next to large depth buffers, one-row stages for
    wide embeddings, a whole patch row in one stage for narrow ones, fp16 tokens) and the ones it must hand to k_fused
    (patch 54 / 63: the depth buffers alone exceed shared memory): the result equals the operator-sequence kernel bit for bit,
    for one frame (fewer patch rows than SMs) and for enough frames that every CTA loops."""
    fe = FrontEnd(device=DEV, patch_size=patch, embedding_size=E, dtype=dtype)
    fe_exact = FrontEnd(device=DEV, patch_size=patch, embedding_size=E, dtype=dtype)
    fe_exact.exact_pooling = True
    for n_frames in (1, 40):
        sc = synthetic.make_scene(60 + n_frames, n_frames, n_total=60, device=DEV)
        tok = synthetic.make_tokens(n_frames, n_frames, fe.tokens_per_frame, E).to(DEV).to(dtype)
        a = fe(sc.depth, sc.poses, sc.intrinsics, tokens=tok)
        b = fe_exact(sc.depth, sc.poses, sc.intrinsics, tokens=tok)
        assert a.dtype == dtype and torch.equal(a.view(torch.int16), b.view(torch.int16)), (patch, E, n_frames)


def test_mc_dense_voxeliser_equals_the_hash_path(monkeypatch):
    """k_mc_voxelize_dense (cell bitset of the universe's box in shared memory) against k_mc_voxelize alone (V3D_MC_DENSE=0:
    queue + hash probes): identical coverage bitsets for full-size scenes, a universe with duplicates, a universe that is
    too large for the cell bitset (hash path for that scene only), one non-rigid frame (hash path for that frame only) and
    an empty universe; a universe that covers only the middle of the room (pixels outside the box on every side)."""
    scenes = []
    for s in range(5):
        seed = 720 + s
        n = 40 if s < 3 else 7
        sc = synthetic.make_scene(seed, n, n_total=300, frame_indices=np.arange(0, 2 * n, 2), device=DEV)
        pc = synthetic.mesh_voxels(sc.room, sc.boxes, seed)
        poses = sc.poses.clone()
        if s == 0:  # the universe is a slab of the room: many pixels fall outside its box, below and above in x, y and z
            lo, hi = np.percentile(pc, 30, axis=0), np.percentile(pc, 75, axis=0)
            pc = np.ascontiguousarray(pc[(pc[:, 0] >= lo[0]) & (pc[:, 0] <= hi[0]) & (pc[:, 1] <= hi[1]) & (pc[:, 2] >= 2) & (pc[:, 2] <= pc[:, 2].max() - 2)])
            assert len(pc) > 500
        if s == 1:
            pc = np.concatenate([pc, pc[:300], pc[100:200]])  # duplicates own one bit
        if s == 2:
            pc = np.concatenate([pc, np.array([[3000, 0, 0], [-3000, 5, 1]], dtype=pc.dtype)])  # box of 6001 cells in x
            poses[3, 3, :] = torch.tensor([0.0, 0.0, 1e-3, 1.0], device=DEV)  # projective frame
        if s == 3:
            poses[2, 3, :] = torch.tensor([0.0, 1e-3, 0.0, 1.0], device=DEV)  # one frame for the hash kernel
            # frames far outside the box (4096 cells and more): not for the dense kernel's combined x / y box test
            poses[4, 1, 3] += 3276.8 + 409.6
            poses[5, 0, 3] -= 32768 * 0.1
            poses[6, :3, 3] += torch.tensor([409.6, 819.2, -1638.4], device=DEV)
        if s == 4:
            pc = pc[:0]
        scenes.append({"depth": sc.depth, "poses": poses, "intrinsics": sc.intrinsics,
                       "pc_voxels": torch.from_numpy(np.ascontiguousarray(pc)).to(DEV), "seed": seed})
    sampler = mc_sampling.MCSampler(device=DEV)
    plan = sampler.plan(scenes)
    monkeypatch.setenv("V3D_MC_DENSE", "0")
    sampler.run(plan, "voxelize")
    ref = [sampler.bitsets(plan, i).copy() for i in range(len(scenes))]
    res_ref = sampler.decode(sampler.run(plan))
    monkeypatch.setenv("V3D_MC_DENSE", "1")
    sampler.run(plan, "voxelize")
    for i in range(len(scenes)):
        got = sampler.bitsets(plan, i)
        if i == 1:  # which of the duplicates owns the bit is the hash build's choice: compare per distinct voxel
            pc = scenes[i]["pc_voxels"].cpu().numpy()
            _, inv = np.unique(pc, axis=0, return_inverse=True)
            for f in range(got.shape[0]):
                a = np.unpackbits(got[f].view(np.uint8), bitorder="little")[: len(pc)]
                b = np.unpackbits(ref[i][f].view(np.uint8), bitorder="little")[: len(pc)]
                np.testing.assert_array_equal(np.bincount(inv.ravel(), a), np.bincount(inv.ravel(), b))
                assert np.bincount(inv.ravel(), a).max() <= 1
        else:
            np.testing.assert_array_equal(got, ref[i])
        assert ref[i].any() or i == 4
    assert sampler.decode(sampler.run(plan)) == res_ref

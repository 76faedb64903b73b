"""CPU: pin the oracle (oracle/v3d_oracle.c via oracle/oracle.py) against outputs of the
unmodified reference (tests/golden/*, written by oracle/gen_golden.py)."""
import hashlib
import json
import os
import random

import numpy as np
import pytest
import torch

import fixtures
from oracle import oracle as O
from video_3d_llm_b200 import synthetic


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ulp_diff(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7FFFFFFF), a)
    b = np.where(b < 0, -(b & 0x7FFFFFFF), b)
    return np.abs(a - b)


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "frontend.npz"))


@pytest.fixture(scope="module")
def scene0(G):
    sc = synthetic.make_scene(0, 2, n_total=30, H=480, W=640, device="cpu")
    assert sha(sc.depth_u16()) == str(G["in_depth_sha"]), "synthetic input drifted from the golden run"
    np.testing.assert_array_equal(sc.poses.numpy(), G["in_poses"])
    np.testing.assert_array_equal(sc.intrinsics.numpy(), G["in_intr"])
    return sc


@pytest.fixture(scope="module")
def stages(scene0):
    return O.frontend(scene0.depth_u16(), scene0.poses.numpy(), scene0.intrinsics.numpy(), patch_size=27,
                      embedding_size=3584)


def test_unproject_bit_exact(G, scene0, stages):
    w = stages["world"]
    np.testing.assert_array_equal(w[:, ::37, ::41], G["world_sample"])
    assert sha(w) == str(G["world_sha"])
    # every depth dtype the boundary accepts gives the same bits
    d = scene0.depth_u16()
    for dd in (d.astype(np.int32), d.astype(np.float32)):
        assert sha(O.unproject(scene0.intrinsics.numpy(), scene0.poses.numpy(), dd)) == str(G["world_sha"])


def test_unproject_vs_float64_scalar_twin(G, scene0, stages):
    w = stages["world"]
    for f in range(2):
        for k, (u, v) in enumerate(G["uvd_px"]):
            np.testing.assert_allclose(w[f, v, u], G["uvd_f64"][f, k], rtol=1e-5, atol=1e-5)
            mine = O.unproject_uvd_f64(u, v, scene0.depth_u16()[f, v, u], scene0.intrinsics[f].numpy(), scene0.poses[f].numpy())
            np.testing.assert_allclose(mine, G["uvd_f64"][f, k], rtol=1e-12, atol=1e-12)


def test_zero_depth_maps_to_camera_centre(scene0, stages):
    d = scene0.depth_u16()
    f, r, c = np.argwhere(d == 0)[0]
    np.testing.assert_array_equal(stages["world"][f, r, c], scene0.poses.numpy()[f, :3, 3])


def test_boundary(G, stages):
    np.testing.assert_array_equal(O.boundary(stages["world"]), G["boundry"])


@pytest.mark.parametrize("hw", [(480, 640), (120, 160), (60, 80), (968, 1296)])
@pytest.mark.parametrize("strat", ["center_crop", "resize"])
def test_resample_maps_match_cv2(G, hw, strat):
    rm, cm = O.resample_maps(hw[0], hw[1], 384, strat)
    np.testing.assert_array_equal(rm, G[f"rowmap_{hw[0]}x{hw[1]}_{strat}"])
    np.testing.assert_array_equal(cm, G[f"colmap_{hw[0]}x{hw[1]}_{strat}"])


def test_default_geometry_is_floor_1p25():
    rm, cm = O.resample_maps(480, 640)
    np.testing.assert_array_equal(rm, np.floor(1.25 * np.arange(384)).astype(np.int32))
    np.testing.assert_array_equal(cm, np.floor(1.25 * (np.arange(384) + 64)).astype(np.int32))


def test_resize_crop_bit_exact(G, stages):
    assert sha(stages["coords"]) == str(G["coords_sha"])
    rm, cm = O.resample_maps(480, 640, 384, "resize")
    assert sha(O.resample(stages["world"], rm, cm)) == str(G["coords_resize_sha"])


def test_pooling_bit_exact(G, stages):
    c = stages["coords"]
    np.testing.assert_array_equal(O.average_coordinate_in_patch(c, 27), G["avg27"])
    np.testing.assert_array_equal(O.average_coordinate_in_patch(c, 14), G["avg14"])
    np.testing.assert_array_equal(O.minmax_coordinate_in_patch(c, 27), G["minmax27"])
    np.testing.assert_array_equal(O.minmax_coordinate_in_patch(c, 14), G["minmax14"])
    for n in (9, 5, 1):
        np.testing.assert_array_equal(O.sample_n_points(c, n), G[f"sample{n}"])


def test_discretise_bit_exact(G):
    np.testing.assert_array_equal(O.discrete_coords(G["avg27"]), G["disc27"])
    np.testing.assert_array_equal(O.discrete_coords(G["avg14"]), G["disc14"])
    np.testing.assert_array_equal(O.discrete_coords(G["sample9"]), G["disc_sample9"])
    np.testing.assert_array_equal(O.discrete_coords(G["minmax27"]), G["disc_minmax27"])
    np.testing.assert_array_equal(O.discrete_coords(G["disc_stress_in"]), G["disc_stress_out"])


@pytest.mark.parametrize("nf", [1194, 384, 385, 132, 597])
def test_dim_t_table(G, nf):
    np.testing.assert_array_equal(O.dim_t_table(nf), G[f"dim_t_{nf}"])


def _pe_close(mine, ref, max_ulp_abs=4e-7):
    # sin/cos come from different libms (glibc vs SLEEF): allow a few ulp of |value|<=1
    assert mine.shape == ref.shape
    assert np.max(np.abs(mine.astype(np.float64) - ref.astype(np.float64))) <= max_ulp_abs
    # after bf16 rounding at most a handful of elements may land on the other side of a tie
    mb, rb = O.f32_to_bf16_bits(mine), O.f32_to_bf16_bits(ref)
    assert np.mean(mb != rb) < 1e-4


def test_pe_sine3d(G):
    d27 = G["disc27"].reshape(2, 196, 3)
    d14 = G["disc14"].reshape(2, 729, 3)
    _pe_close(O.pe_sine3d(d27, 3584)[:, ::13], G["pe3584_sel"])
    _pe_close(O.pe_sine3d(d14, 1152)[:, ::37], G["pe1152_sel"])
    _pe_close(O.pe_sine3d(d14, 1155)[:, ::37], G["pe1155_sel"])  # odd nf: last channel of each axis is a sin
    _pe_close(O.pe_sine3d(G["avg27"].reshape(2, 196, 3), 3584)[:, ::13], G["pe3584_cont_sel"])
    _pe_close(O.pe_sine3d(G["disc_sample9"].reshape(2, 196, 9, 3), 3584, n_points=9)[:, ::13], G["pe3584_n9_sel"])
    _pe_close(O.pe_sine3d(G["disc_minmax27"].reshape(2, 196, 2, 3), 3584, n_points=2)[:, ::13], G["pe3584_n2_sel"])
    # zero tail: 3584 = 3*1194 + 2
    assert np.all(O.pe_sine3d(d27, 3584)[..., 3582:] == 0)


def test_pe_table_form(G):
    """for discretised coordinates the PE is a 3-row gather from a 301-row table."""
    idx = np.repeat(np.arange(301, dtype=np.float32)[None, :, None], 3, 2)
    t = O.pe_sine3d(idx, 3584)[0, :, :1194]
    _pe_close(t[::7], G["pe_table_3584_rows7"])
    d27 = G["disc27"].reshape(-1, 3).astype(np.int64)
    gathered = np.concatenate([t[d27[:, 0]], t[d27[:, 1]], t[d27[:, 2]], np.zeros((len(d27), 2), np.float32)], 1)
    np.testing.assert_array_equal(gathered.reshape(2, 196, 3584), O.pe_sine3d(G["disc27"].reshape(2, 196, 3), 3584))


def test_bf16_rounding_matches_torch():
    x = np.random.default_rng(0).standard_normal(100000).astype(np.float32)
    x[:4] = [np.nan, np.inf, -np.inf, 0.0]
    mine = O.f32_to_bf16_bits(x)
    ref = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    np.testing.assert_array_equal(mine[1:], ref[1:])
    assert np.isnan(O.bf16_bits_to_f32(mine[:1]))[0]


def test_add_bf16(G):
    tokens = synthetic.make_tokens(0, 2, 196, 3584)
    tb = tokens.view(torch.int16).numpy().view(np.uint16)
    assert sha(tb) == str(G["in_tokens_sha"])
    pe = O.pe_sine3d(G["disc27"].reshape(2, 196, 3), 3584)
    out = O.add_pe_bf16(tb, pe)
    ref_sel = O.f32_to_bf16_bits(G["added3584_sel"])
    assert np.mean(out[:, ::13] != ref_sel) < 1e-4
    # exact when fed the reference's own fp32 PE rows
    out_sel = O.add_pe_bf16(tb[:, ::13], G["pe3584_sel"])
    np.testing.assert_array_equal(out_sel, ref_sel)


def test_pe_mlp(G):
    out = O.pe_mlp(G["avg14"].reshape(2, 729, 3), G["mlp_mlp_0_weight"], G["mlp_mlp_0_bias"], G["mlp_mlp_1_weight"],
                   G["mlp_mlp_1_bias"], G["mlp_mlp_3_weight"], G["mlp_mlp_3_bias"])
    np.testing.assert_allclose(out[:, ::37], G["mlp_out_sel"], rtol=2e-4, atol=2e-5)


# ---------------------------------------------------------------------------------------------
# MC sampler and the host-side sampling logic
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def D(golden_dir):
    with open(os.path.join(golden_dir, "dataset.json")) as f:
        return json.load(f)


def test_mt19937_matches_cpython(golden_dir):
    with open(os.path.join(golden_dir, "mt19937.json")) as f:
        kat = json.load(f)
    for seed, e in kat.items():
        np.testing.assert_array_equal(O.mc_randbelow_seq(int(seed), e["ns"]), np.array(e["picks"], np.uint32))
    # and live against this interpreter's random module, incl. 1-element choices consuming words
    random.seed(77)
    ns = [1, 1, 1, 5, 1, 9, 150]
    assert [random.choice(range(n)) for n in ns] == list(O.mc_randbelow_seq(77, ns))


def _scene_inputs(D, name):
    meta = D["scenes"][name]
    sc = synthetic.make_scene(meta["seed"], meta["n_total"], n_total=meta["n_total"], H=meta["H"], W=meta["W"], device="cpu")
    assert sha(sc.depth_u16()) == meta["depth_sha"], "synthetic input drifted from the golden run"
    pc = synthetic.mesh_voxels(sc.room, sc.boxes, meta["seed"])
    assert sha(pc) == meta["pc_sha"]
    cand = fixtures.mc_candidates(meta["n_total"])
    return sc, pc, cand, fixtures.disk_poses(sc)


@pytest.mark.parametrize("name", ["scene0001_00", "scene0002_00", "scene0003_00"])
def test_mc_sampler_bit_exact(D, name):
    sc, pc, cand, poses = _scene_inputs(D, name)
    gold = next(s for s in D["select_frames"] if s["video_id"].endswith(name))
    order, gains, n_all, n_sel, bs = O.mc_sample_scene(sc.depth_u16()[cand], poses[cand], sc.intrinsics.numpy()[cand], pc,
                                                        0.1, 32, D["scenes"][name]["seed"])
    gold_idx = [int(f.split("/")[-1].split(".")[0]) // 10 for f in gold["frame_files"]]
    assert [cand[i] for i in order] == gold_idx  # including the seeded tie-breaks
    assert list(gains) == gold["voxel_nums"]
    assert n_all == gold["num_all_voxels"] and n_sel == gold["num_select_voxels"]
    if name != "scene0003_00":  # small enough for the pure-Python set version
        ids = O.mc_voxel_ids(O.unproject(sc.intrinsics.numpy()[cand], poses[cand], sc.depth_u16()[cand]))
        o2, g2, a2, s2 = O.mc_python_sets(ids, pc, 32, D["scenes"][name]["seed"])
        assert list(o2) == list(order) and list(g2) == list(gains) and (a2, s2) == (n_all, n_sel)


@pytest.mark.parametrize("strat", ["mc", "mc-ratio90", "mc-ratio95"])
def test_mc_prefix_cut(D, strat):
    for s in D["select_frames"]:
        name = s["video_id"].split("/")[-1]
        for ub in (32, 8, 3):
            mine = O.mc_prefix_cut(s["frame_files"], s["voxel_nums"], s["num_all_voxels"], strat, ub)
            assert mine == D["video_processor"][strat][name][f"files_ub{ub}"]


# ---------------------------------------------------------------------------------------------
# the torch-CPU port timed by bench.py as the CPU baseline is pinned too
# ---------------------------------------------------------------------------------------------
def test_torch_port_bit_exact(G, scene0):
    from oracle import torch_port as TP

    depth = torch.from_numpy(scene0.depth_u16().astype(np.int32))
    tokens = synthetic.make_tokens(0, 2, 196, 3584)
    out, pooled, disc = TP.frontend(depth, scene0.poses, scene0.intrinsics, 27, 3584, tokens=tokens)
    world = TP.unproject(scene0.intrinsics, scene0.poses, depth.float())
    assert sha(world.numpy()) == str(G["world_sha"])
    np.testing.assert_array_equal(TP.boundary(world).numpy(), G["boundry"])
    assert sha(TP.resize_center_crop(world).numpy()) == str(G["coords_sha"])
    np.testing.assert_array_equal(pooled.numpy(), G["avg27"])
    np.testing.assert_array_equal(disc.numpy(), G["disc27"])
    assert sha(out.view(torch.int16).numpy()) == str(G["added3584_bf16_sha"])
    pe, pooled14, disc14 = TP.frontend(depth, scene0.poses, scene0.intrinsics, 14, 1152)
    np.testing.assert_array_equal(pooled14.numpy(), G["avg14"])
    assert sha(pe.view(torch.int16).numpy()) == str(G["pe1152_bf16_sha"])


# ---------------------------------------------------------------------------------------------
# object proposals (section 8 f1)
# ---------------------------------------------------------------------------------------------
def test_object_patch_masks_and_features(golden_dir, stages):
    GO = np.load(os.path.join(golden_dir, "objects.npz"))
    assert sha(stages["coords"]) == str(GO["coords_sha"])
    for kind in ("patch14", "patch27"):
        shape = tuple(GO[f"masks_{kind}_shape"])
        ref = np.unpackbits(GO[f"masks_{kind}"])[: int(np.prod(shape))].reshape(shape).astype(bool)
        mine = O.object_patch_masks(stages["coords"], GO["boxes"], kind)
        np.testing.assert_array_equal(mine, ref)
        assert ref.reshape(shape[0], -1).sum(1).tolist() == GO[f"masks_{kind}_count"].tolist()
    assert GO["masks_patch27_count"].max() > 50 and (GO["masks_patch27_count"] == 0).any()  # the fixture exercises both
    tokens = synthetic.make_tokens(0, 2, 196, 256).view(torch.int16).numpy().view(np.uint16).reshape(-1, 256)
    feat = O.object_features(tokens, O.object_patch_masks(stages["coords"], GO["boxes"], "patch27"))
    np.testing.assert_allclose(O.bf16_bits_to_f32(feat), GO["objfeat_patch27"], rtol=0, atol=2.0 ** -7)


# ---------------------------------------------------------------------------------------------
# projector tail: the torch port == the reference's get_2dPool / PE / add_token_per_grid
# ---------------------------------------------------------------------------------------------
TAIL_CASES = ["bilinear_1152", "bilinear_3584", "average_1152", "max_1152", "bilinear_16_264"]


@pytest.mark.parametrize("tag", TAIL_CASES)
def test_torch_port_projector_tail_matches_reference(golden_dir, tag):
    import hashlib

    import torch

    from oracle import torch_port as TP
    from video_3d_llm_b200 import synthetic

    G = np.load(os.path.join(golden_dir, "tail.npz"))
    V, side, E, Go = (int(v) for v in G[f"{tag}_meta"])
    tokens = synthetic.make_tokens(70, V, side * side, E)
    coords = torch.from_numpy(G[f"{tag}_coords"]).to(torch.bfloat16)
    newline = torch.from_numpy(G[f"{tag}_newline"]).view(torch.bfloat16)
    mode = tag.split("_")[0]
    pooled = TP.get_2dpool(tokens, 2, mode)
    out = TP.projector_tail(tokens, coords, E, newline, 2, mode)
    assert out.shape == (V * Go * (Go + 1), E)
    sha = lambda t: hashlib.sha256(np.ascontiguousarray(t.contiguous().view(torch.int16).numpy()).tobytes()).digest()
    assert sha(pooled) == G[f"{tag}_pooled_sha"].tobytes()
    assert sha(out) == G[f"{tag}_out_sha"].tobytes()
    np.testing.assert_array_equal(out.contiguous().view(torch.int16).numpy()[::7, ::5], G[f"{tag}_out_sample"])

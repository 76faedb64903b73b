"""CPU: host-side logic of the mirror (no kernels): geometry maps, option-string parsing, frame sampling
against the reference's own outputs (tests/golden/dataset.json), lazy handles, loud failure without CUDA."""
import json
import os

import numpy as np
import pytest
import torch

import fixtures
from video_3d_llm_b200 import coords, geometry, synthetic, video_utils as VU
from video_3d_llm_b200.position_encoding import PositionEmbeddingMLP, PositionEmbeddingSine3D, sine3d_dim_t


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "frontend.npz"))


@pytest.mark.parametrize("hw", [(480, 640), (120, 160), (60, 80), (968, 1296)])
@pytest.mark.parametrize("strat", ["center_crop", "resize"])
def test_geometry_maps_match_cv2(G, hw, strat):
    rows, cols = geometry.resample_maps_host(hw[0], hw[1], 384, strat)
    np.testing.assert_array_equal(np.array(rows), G[f"rowmap_{hw[0]}x{hw[1]}_{strat}"])
    np.testing.assert_array_equal(np.array(cols), G[f"colmap_{hw[0]}x{hw[1]}_{strat}"])
    with pytest.raises(ValueError):
        geometry.resample_maps_host(480, 640, 384, "letterbox")


@pytest.mark.parametrize("nf", [1194, 384, 385, 132, 597])
def test_dim_t_is_the_reference_table(G, nf):
    np.testing.assert_array_equal(sine3d_dim_t(nf).numpy(), G[f"dim_t_{nf}"])


def test_option_string_semantics():
    """Substring matching of world_position_embedding_type (llava_arch.py:51-63, 395-429)."""
    p = coords.parse_world_position_embedding_type
    assert p("avg-discrete-sin3d") == {"pool": "avg", "n_points": 1, "discrete": True, "sin3d": True, "mlp": False, "mrope": False}
    assert p("sample9-sin3d")["n_points"] == 9 and p("sample5-mlp")["n_points"] == 5 and p("sample1-discrete")["n_points"] == 1
    assert p("minmax-discrete-sin3d")["n_points"] == 2 and p("minmax-discrete-sin3d")["pool"] == "minmax"
    assert p("avg-discrete-mrope")["mrope"] and not p("avg-discrete-mrope")["sin3d"]
    assert p("none")["pool"] is None


def test_modules_keep_reference_surface():
    m = PositionEmbeddingSine3D(3584, n_points=9)
    assert (m.embedding_size, m.temperature, m.n_points) == (3584, 10000, 9) and len(list(m.parameters())) == 0
    mlp = PositionEmbeddingMLP(1152)
    assert sorted(mlp.state_dict().keys()) == ["mlp.0.bias", "mlp.0.weight", "mlp.1.bias", "mlp.1.weight", "mlp.3.bias", "mlp.3.weight"]
    assert mlp.mlp[0].weight.shape == (512, 3) and mlp.mlp[3].weight.shape == (1152, 512)


def test_product_fails_loudly_without_cuda():
    x = torch.zeros(1, 4, 3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        PositionEmbeddingSine3D(96)(x)
    with pytest.raises(RuntimeError, match="no CPU path"):
        coords.average_coordinate_in_patch(torch.zeros(1, 384, 384, 3))
    with pytest.raises(RuntimeError, match="no CPU path"):
        coords.discrete_coords(x)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            VU.unproject(torch.eye(4)[None], torch.eye(4)[None], torch.zeros(1, 8, 8))


def _vp(root, strategy, **kw):
    return VU.VideoProcessor(video_folder=os.path.join(root, "data"), annotation_dir=os.path.join(root, "data", "embodiedscan"),
                             voxel_size=0.1, min_xyz_range=[-15, -15, -5], max_xyz_range=[15, 15, 5],
                             frame_sampling_strategy=strategy, metadata_dir=os.path.join(root, "data", "metadata"), **kw)


def _rel(files, root):
    return [os.path.relpath(f, root) if os.path.isabs(f) else f for f in files]


@pytest.mark.parametrize("strategy", ["uniform", "mc", "mc-ratio90", "mc-ratio95", "mc-ratio95-norm"])
def test_frame_sampling_matches_reference(dataset, D, strategy):
    vp = _vp(dataset, strategy)
    gold = D["video_processor"][strategy]
    for name, e in gold.items():
        vid = f"scannet/{name}"
        if "mc" in strategy:
            for ub in (32, 8, 3):
                assert vp.sample_frame_files_mc(vid, frames_upbound=ub) == e[f"files_ub{ub}"]
            assert vp.pc_min[name].tolist() == e["pc_min"] and vp.pc_max[name].tolist() == e["pc_max"]
        else:
            assert _rel(vp.sample_frame_files(vid, force_sample=True, frames_upbound=8), dataset) == e["files_force8"]
            assert _rel(vp.sample_frame_files(vid, force_sample=False, frames_upbound=8), dataset) == e["files_default"]


def test_discrete_point_matches_reference(dataset, D):
    vp = _vp(dataset, "uniform")
    assert vp.discrete_point(D["discrete_point"]["in"]) == D["discrete_point"]["out"]
    vp2 = VU.VideoProcessor(video_folder=os.path.join(dataset, "data"), annotation_dir=os.path.join(dataset, "data", "embodiedscan"),
                            voxel_size=0.1, metadata_dir=os.path.join(dataset, "data", "metadata"))
    assert vp2.discrete_point(D["discrete_point_norange"]["in"]) == D["discrete_point_norange"]["out"]


def test_lazy_preprocess_and_merge_without_touching_cuda(dataset, D):
    """DataLoader workers must not need CUDA: preprocess() returns raw depth/pose/intrinsics behind a handle."""
    vp = _vp(dataset, "uniform", lazy=True)
    ip = fixtures.FakeImageProcessor(384)
    name = "scene0003_00"
    d = vp.process_3d_video(f"scannet/{name}", ip, force_sample=True, frames_upbound=3)
    gold = D["video_processor"]["uniform"][name]
    wc = d["world_coords"]
    assert isinstance(wc, VU.LazyWorldCoords) and list(wc.shape) == gold["world_coords_shape"]
    assert wc.depth.dtype == torch.int16 and wc.depth.shape == (3, 480, 640) and not wc.depth.is_cuda
    assert d["video_size"] == gold["video_size"] and list(d["images"].shape) == gold["images_shape"]
    assert d["objects"].tolist() == gold["objects"]
    # poses are float32(axis_align @ loadtxt(pose)) exactly as the reference builds them
    sc = synthetic.make_scene(3, 4, n_total=4, device="cpu")
    idx = np.linspace(0, 3, 3, dtype=int)
    np.testing.assert_array_equal(wc.poses.numpy(), fixtures.disk_poses(sc)[idx])
    np.testing.assert_array_equal(wc.depth.numpy().view(np.uint16), sc.depth_u16()[idx])
    d["box_input"] = None
    m = VU.merge_video_dict([d, dict(d)])
    assert sorted(m.keys()) == gold["merged_keys"]
    assert list(m["world_coords"].shape) == [2] + gold["world_coords_shape"] and len(m["box_input"]) == 0
    scenes = list(m["world_coords"])
    assert len(scenes) == 2 and list(scenes[0].shape) == gold["world_coords_shape"]
    # what llava_arch.py:387 does to the tensor is accepted without evaluating anything
    xyz_min = m["world_coords"].view(2, -1, 3).min(dim=1)[0]
    assert xyz_min[0] is not None
    half = m["world_coords"].half()
    assert half.dtype == torch.float16 and half.depth is m["world_coords"].depth
    d2 = dict(d)
    d2["box_input"] = [1.0, 2.0, 3.0]
    assert VU.merge_video_dict([d2])["box_input"].shape == (1, 3)


def test_mc_ratio_cut_matches_the_reference_prefix_rule():
    """mc_sampling.ratio_cut (candidate indices) against the oracle's restatement of sample_frame_files_mc
    (llava/video_utils.py:131-159) on file names, incl. exact-threshold and no-cut cases."""
    from oracle import oracle as O
    from video_3d_llm_b200 import mc_sampling

    rng = np.random.default_rng(5)
    for trial in range(50):
        n = int(rng.integers(1, 33))
        order = rng.permutation(150)[:n].tolist()
        gains = sorted(rng.integers(0, 400, n).tolist(), reverse=True)
        n_all = int(sum(gains) + rng.integers(0, 50))
        if trial % 7 == 0:
            n_all = int(round(sum(gains[: max(1, n // 2)]) / 0.95))  # lands on / next to the threshold
        rec = {"order": order, "gains": gains, "n_all": n_all, "n_sel": sum(gains)}
        files = [f"data/scannet/posed_images/s/{10 * f:05d}.jpg" for f in order]
        for strategy in ("mc-ratio95", "mc-ratio90-norm", "mc", "mc-online-ratio95"):
            for ub in (32, 8, 3):
                got = mc_sampling.ratio_cut(rec, strategy, ub)
                ref = O.mc_prefix_cut(files, gains, n_all, strategy, ub)
                assert [f"data/scannet/posed_images/s/{10 * f:05d}.jpg" for f in got] == ref

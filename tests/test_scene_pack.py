"""Packed per-scene binary: byte-identical inputs to the PNG/txt path of the reference
(llava/video_utils.py:212-230), format checks, error behaviour.  CPU only."""
import os
import struct

import numpy as np
import pytest
import torch

from video_3d_llm_b200 import scene_pack
from video_3d_llm_b200 import video_utils as VU


def _vp(root, **kw):
    return VU.VideoProcessor(video_folder=os.path.join(root, "data"), annotation_dir=os.path.join(root, "data", "embodiedscan"),
                             voxel_size=0.1, min_xyz_range=[-15, -15, -5], max_xyz_range=[15, 15, 5],
                             frame_sampling_strategy="uniform", metadata_dir=os.path.join(root, "data", "metadata"), **kw)


@pytest.fixture(scope="module")
def packed(dataset, tmp_path_factory):
    out = str(tmp_path_factory.mktemp("packs"))
    vp = _vp(dataset)
    for video_id in vp.scene:
        scene_pack.pack_scene(vp, video_id, scene_pack.pack_path(out, video_id))
    return out


def test_pack_equals_png_txt_path(dataset, packed):
    vp_files, vp_pack = _vp(dataset), _vp(dataset, pack_dir=packed)
    for video_id, info in vp_files.scene.items():
        files = [os.path.join(vp_files.video_folder, img["img_path"]) for img in info["images"]]
        for pick in (files, files[::3], files[::-2][:5], files[:1]):
            a = vp_files.load_frames(video_id, pick)
            b = vp_pack.load_frames(video_id, pick)
            for x, y in zip(a, b):
                assert x.dtype == y.dtype and x.shape == y.shape
                assert torch.equal(x, y)
        assert vp_pack._pack(video_id) is not None


def test_pack_layout_and_header(dataset, packed):
    vp = _vp(dataset)
    video_id = sorted(vp.scene)[0]
    path = scene_pack.pack_path(packed, video_id)
    raw = open(path, "rb").read()
    magic, ver, n, H, W, code, _, off_ids, off_poses, off_intr, off_depth, total = struct.unpack_from("<8sIIIIIIQQQQQ", raw)
    assert magic == b"V3DPACK1" and ver == 2 and code == 0 and total == len(raw)
    assert struct.unpack_from("<QII", raw, struct.calcsize("<8sIIIIIIQQQQQ")) == (0, 0, 0)  # no resampled stack in this pack
    assert n == len(vp.scene[video_id]["images"]) and off_depth % 4096 == 0 and off_ids == 4096
    assert total == off_depth + n * H * W * 2
    pk = scene_pack.ScenePack(path)
    assert pk.ids.tolist() == [scene_pack.frame_number(i["img_path"]) for i in vp.scene[video_id]["images"]]
    d, p, k = pk.load(pk.ids[:2], pin=False)
    assert d.dtype == torch.int16 and tuple(d.shape) == (2, H, W) and tuple(p.shape) == (2, 4, 4) and tuple(k.shape) == (2, 4, 4)


def test_pack_errors(dataset, packed, tmp_path):
    vp = _vp(dataset)
    video_id = sorted(vp.scene)[0]
    pk = scene_pack.ScenePack(scene_pack.pack_path(packed, video_id))
    with pytest.raises(KeyError):
        pk.load([123457])
    bad = tmp_path / "bad.v3dpack"
    bad.write_bytes(b"NOTAPACK" + bytes(4096))
    with pytest.raises(ValueError):
        scene_pack.ScenePack(str(bad))
    raw = open(scene_pack.pack_path(packed, video_id), "rb").read()
    trunc = tmp_path / "trunc.v3dpack"
    trunc.write_bytes(raw[:-10])
    with pytest.raises(ValueError):
        scene_pack.ScenePack(str(trunc))
    with pytest.raises(ValueError):
        scene_pack.write_pack(str(tmp_path / "x.v3dpack"), [0], np.eye(4, dtype=np.float32)[None], np.eye(4, dtype=np.float32),
                              np.zeros((1, 4, 4), np.float32))
    # a scene without a pack falls back to the files
    vp2 = _vp(dataset, pack_dir=str(tmp_path))
    files = [os.path.join(vp2.video_folder, img["img_path"]) for img in vp2.scene[video_id]["images"]][:2]
    assert torch.equal(vp2.load_frames(video_id, files)[0], vp.load_frames(video_id, files)[0])


def test_int32_depth_pack_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    depth = rng.integers(0, 70000, (3, 6, 8)).astype(np.int32)
    poses = rng.normal(size=(3, 4, 4)).astype(np.float32)
    intr = rng.normal(size=(4, 4)).astype(np.float32)
    path = str(tmp_path / "s.v3dpack")
    scene_pack.write_pack(path, [30, 10, 20], poses, intr, depth)
    d, p, k = scene_pack.ScenePack(path).load([10, 30])
    assert d.dtype == torch.int32 and np.array_equal(d.numpy(), depth[[1, 0]])
    assert np.array_equal(p.numpy(), poses[[1, 0]]) and np.array_equal(k.numpy(), np.repeat(intr[None], 2, 0))


def test_windowed_load_keeps_exactly_the_samples_the_front_end_reads(dataset, packed):
    vp = _vp(dataset)
    video_id = sorted(vp.scene)[0]
    pk = scene_pack.ScenePack(scene_pack.pack_path(packed, video_id))
    if pk.H % 5:
        pytest.skip("fixture frames are not a multiple of 5 rows")
    es = 2
    x_off, width = 16 * ((pk.W // 8) * es // 16), 16 * ((pk.W // 2) * es // 16)
    window = (x_off, width, 5, 4)
    full, _, _ = pk.load(pk.ids[:3])
    win, _, _ = pk.load(pk.ids[:3], window=window)
    assert tuple(win.shape) == (3, pk.H // 5 * 4, width // es)
    ref = full.view(3, pk.H // 5, 5, pk.W)[:, :, :4, x_off // es:(x_off + width) // es].reshape(win.shape)
    assert torch.equal(win, ref)
    with pytest.raises(ValueError):
        pk.load(pk.ids[:1], window=(0, 2 * pk.W * es, 1, 1))


def test_resampled_load_is_the_nearest_resize_and_crop_of_the_depth(dataset, packed):
    from video_3d_llm_b200 import geometry

    vp = _vp(dataset)
    video_id = sorted(vp.scene)[0]
    pk = scene_pack.ScenePack(scene_pack.pack_path(packed, video_id))
    rows, cols = geometry.resample_maps_host(pk.H, pk.W, 384, "center_crop")
    rows, cols = rows[:378], cols[:378]
    full, _, _ = pk.load(pk.ids[:2])
    res, _, _ = pk.load(pk.ids[:2], resample=(rows, cols))
    assert tuple(res.shape) == (2, 378, 378)
    assert torch.equal(res, full[:, list(rows)][:, :, list(cols)])
    with pytest.raises(ValueError):
        pk.load(pk.ids[:1], resample=([0, pk.H], [0]))


def test_stored_resampled_stack_and_lazy_handle(dataset, tmp_path):
    """Version-2 packs can hold the depth already resampled for the model-side front end: VideoProcessor(pack_dir, lazy=True)
    then hands out handles over those samples (one contiguous copy per frame), bit-identical to resampling on the fly."""
    from video_3d_llm_b200 import geometry

    vp = _vp(dataset)
    video_id = "scannet/scene0003_00"  # 480 x 640
    out = scene_pack.pack_scene(vp, video_id, scene_pack.pack_path(str(tmp_path), video_id), resample=(384, "center_crop"))
    pk = scene_pack.ScenePack(out)
    assert pk.resampled_for == (384, "center_crop") and pk.resampled.shape == (pk.n_frames, 384, 384)
    rows, cols = geometry.resample_maps_host(pk.H, pk.W, 384, "center_crop")
    assert np.array_equal(pk.resampled, pk.depth[:, list(rows)][:, :, list(cols)])
    raw = open(out, "rb").read()
    off_res, crop, strat = struct.unpack_from("<QII", raw, struct.calcsize("<8sIIIIIIQQQQQ"))
    assert off_res % 4096 == 0 and (crop, strat) == (384, 0) and len(raw) == off_res + pk.n_frames * 384 * 384 * 2
    d, p, k = pk.load_resampled(pk.ids[1:3])
    assert d.dtype == torch.int16 and tuple(d.shape) == (2, 384, 384) and torch.equal(d, torch.from_numpy(np.array(pk.resampled[1:3]).view(np.int16)))
    ring = torch.zeros((4, 384, 384), dtype=torch.int16)
    pk.load_resampled(pk.ids[:2], out=ring[2:])
    assert torch.equal(ring[2:], torch.from_numpy(np.array(pk.resampled[:2]).view(np.int16))) and not ring[:2].any()
    vp_pack = _vp(dataset, pack_dir=str(tmp_path), lazy=True)
    files = vp_pack.sample_frame_files(video_id, force_sample=True, frames_upbound=3)
    h = vp_pack.world_coords(video_id, files)
    assert isinstance(h, VU.LazyWorldCoords) and h.resampled and h.src_hw == (480, 640) and h.shape == (3, 384, 384, 3)
    full, _, _ = vp_pack.load_frames(video_id, files)
    assert torch.equal(h.depth, full[:, list(rows)][:, :, list(cols)])
    # another crop size / strategy, or lazy=False, falls back to full frames
    assert not vp_pack.world_coords(video_id, files, crop_size=224).resampled
    with pytest.raises(KeyError):
        scene_pack.ScenePack(scene_pack.pack_path(str(tmp_path), video_id)).load_resampled([123456])
    with pytest.raises(ValueError):
        VU.LazyWorldCoords(h.depth, h.poses, h.intrinsics, resampled=True)
    # version-1 files (no trailer) stay readable
    hdr = bytearray(raw[:4096])
    struct.pack_into("<I", hdr, 8, 1)
    hdr[struct.calcsize("<8sIIIIIIQQQQQ"):struct.calcsize("<8sIIIIIIQQQQQ") + 16] = b"\0" * 16
    off_depth = struct.unpack_from("<Q", hdr, 8 + 24 + 24)[0]
    body = raw[4096:off_depth + pk.n_frames * 480 * 640 * 2]
    struct.pack_into("<Q", hdr, 8 + 24 + 32, 4096 + len(body))
    v1 = tmp_path / "v1.v3dpack"
    v1.write_bytes(bytes(hdr) + body)
    pk1 = scene_pack.ScenePack(str(v1))
    assert pk1.resampled is None and np.array_equal(pk1.depth, pk.depth)


def test_scene_batch_loader_batches_and_thread_safety(dataset, tmp_path):
    """SceneBatchLoader (host path): every batch holds exactly the frames pack.load_resampled gives, batches come in order and
    wrap with loop=True, and loader threads racing for the same scene get ONE mapping (a replaced ScenePack would be
    unmapped under a running native gather -- the segfault of round 2)."""
    import threading

    vp = _vp(dataset)
    ids = [v for v in vp.scene if vp.scene[v]["images"]][:2]
    for video_id in ids:
        scene_pack.pack_scene(vp, video_id, scene_pack.pack_path(str(tmp_path), video_id), resample=(384, "center_crop"))
    vp_pack = _vp(dataset, pack_dir=str(tmp_path), lazy=True)
    sizes = {v: scene_pack.ScenePack(scene_pack.pack_path(str(tmp_path), v)) for v in ids}
    ids = [v for v in ids if (sizes[v].H, sizes[v].W) == (sizes[ids[0]].H, sizes[ids[0]].W)]
    samples = [(v, vp_pack.sample_frame_files(v, force_sample=True, frames_upbound=3)) for v in ids] * 2
    # the race: many threads ask for the same scene's pack at once
    got = []
    vp_race = _vp(dataset, pack_dir=str(tmp_path), lazy=True)
    ts = [threading.Thread(target=lambda: got.append(vp_race._pack(ids[0]))) for _ in range(16)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert all(g is got[0] for g in got)
    loader = VU.SceneBatchLoader(vp_pack, samples, batch_scenes=1, threads=3, prefetch=2, loop=True, pin=False)
    it = iter(loader)
    for n in range(2 * len(samples) + 1):
        h = next(it)
        video_id, files = samples[n % len(samples)]
        d, p, k = vp_pack._pack(video_id).load_resampled([scene_pack.frame_number(f) for f in files])
        assert h.batched and h.resampled and tuple(h.depth.shape) == (1, 3, 384, 384)
        assert torch.equal(h.depth[0], d) and torch.equal(h.poses[0], p) and torch.equal(h.intrinsics[0], k)
    loader.close()
    with pytest.raises(ValueError):
        VU.SceneBatchLoader(vp_pack, [], batch_scenes=1)

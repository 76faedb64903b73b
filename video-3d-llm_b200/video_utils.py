"""Mirror of llava/video_utils.py on the CUDA library: `unproject`, `VideoProcessor`,
`merge_video_dict` keep the reference's names, arguments, option-string semantics and outputs.

Two ways to consume the geometry:
  * eager  (lazy=False): `unproject` / `calculate_world_coords` / `preprocess` return real tensors
    like the reference, computed by the CUDA kernels on the current device.
  * lazy   (lazy=True, the default of VideoProcessor): `world_coords` is a `LazyWorldCoords`
    handle over the raw uint16 depth, poses and intrinsics.  DataLoader workers never touch CUDA
    (it cannot be initialised in forked workers) and ship 3x less data; the fused kernel runs where
    the model runs, when `average_coordinate_in_patch` & co. (coords.py) receive the handle.
    Anything else done to the handle materialises the (V,384,384,3) tensor on demand, so code
    written against the reference keeps working.
"""
from __future__ import annotations

import json
import os
import pickle

import numpy as np
import torch
from PIL import Image

from . import cabi, geometry

DEPTH_SCALE = 1000.0


def _device_for(*tensors):
    """The CUDA device of the first CUDA tensor, else the current device; raises without CUDA."""
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.is_cuda:
            return t.device
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: the video-3d-llm_b200 front end has no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def _as_depth(depths: torch.Tensor) -> torch.Tensor:
    """Keep uint16/int16/int32/float32 depth as is; anything else goes through float32 like the
    reference's `.float()` (llava/video_utils.py:230)."""
    if depths.dtype in (torch.int16, torch.int32, torch.float32) or depths.dtype == getattr(torch, "uint16", None):
        return depths.contiguous()
    return depths.to(torch.float32).contiguous()


def unproject(intrinsics, poses, depths, *, row_map=None, col_map=None, clamp=None):
    """llava/video_utils.py:38-68.  intrinsics (V,4,4), poses (V,4,4), depths (V,H,W) ->
    world coordinates (V,H,W,3) float32 on the CUDA device (inputs on the CPU are moved there).
    row_map/col_map (int32 device tensors) gather a resampled/cropped window instead."""
    lib = cabi.load()
    dev = _device_for(depths, poses, intrinsics)
    depths = _as_depth(depths).to(dev, non_blocking=True)
    poses = poses.to(dev, torch.float32).contiguous()
    intrinsics = intrinsics.to(dev, torch.float32).contiguous()
    V, H, W = depths.shape
    if poses.shape != (V, 4, 4) or intrinsics.shape != (V, 4, 4):
        raise ValueError("intrinsics and poses must be (V,4,4)")
    oh, ow = (H, W) if row_map is None else (int(row_map.numel()), int(col_map.numel()))
    out = torch.empty((V, oh, ow, 3), dtype=torch.float32, device=dev)
    lo_hi = None if clamp is None else cabi.host_floats(list(clamp[0]) + list(clamp[1]))
    with torch.cuda.device(dev):
        cabi.check(lib.v3d_unproject(depths.data_ptr(), cabi.depth_code(depths), poses.data_ptr(), intrinsics.data_ptr(), V, H, W,
                                     DEPTH_SCALE, cabi.ptr(row_map), cabi.ptr(col_map), oh, ow, lo_hi, out.data_ptr(),
                                     cabi.stream_ptr(dev)), "v3d_unproject")
    return out


def boundary(intrinsics, poses, depths, clamp=None) -> torch.Tensor:
    """[xmin,xmax,ymin,ymax,zmin,zmax] of the unprojected scene (llava/video_utils.py:273-277)."""
    lib = cabi.load()
    dev = _device_for(depths, poses, intrinsics)
    depths = _as_depth(depths).to(dev)
    poses = poses.to(dev, torch.float32).contiguous()
    intrinsics = intrinsics.to(dev, torch.float32).contiguous()
    V, H, W = depths.shape
    out = torch.empty(6, dtype=torch.float32, device=dev)
    lo_hi = None if clamp is None else cabi.host_floats(list(clamp[0]) + list(clamp[1]))
    with torch.cuda.device(dev):
        cabi.check(lib.v3d_boundary(depths.data_ptr(), cabi.depth_code(depths), poses.data_ptr(), intrinsics.data_ptr(), V, H, W,
                                    DEPTH_SCALE, lo_hi, out.data_ptr(), cabi.stream_ptr(dev)), "v3d_boundary")
    return out


class _LazyMin:
    """Result of `lazy.view(B,-1,3).min(dim=1)`: llava_arch.py:387 computes it and never uses the value
    (discrete_coords ignores its xyz_min argument), so it is only evaluated if somebody looks."""

    def __init__(self, owner):
        self._owner = owner
        self._val = None

    def _get(self):
        if self._val is None:
            full = self._owner.materialize()
            self._val = full.reshape(self._owner.batch_size, -1, 3).min(dim=1)
        return self._val

    def __getitem__(self, i):
        return _LazyMinItem(self, i)

    def __iter__(self):
        return iter(self._get())


class _LazyMinItem:
    def __init__(self, parent, i):
        self._parent, self._i = parent, i

    def _get(self):
        return self._parent._get()[self._i]

    def __getitem__(self, j):
        return _LazyMinItem(self, j)

    def __getattr__(self, name):
        return getattr(self._get(), name)


class _LazyFlat:
    def __init__(self, owner):
        self._owner = owner

    def min(self, dim=None, **kw):
        if dim == 1:
            return _LazyMin(self._owner)
        return self._owner.materialize().reshape(self._owner.batch_size, -1, 3).min(dim=dim, **kw)

    def __getattr__(self, name):
        return getattr(self._owner.materialize().reshape(self._owner.batch_size, -1, 3), name)


class LazyWorldCoords:
    """Stand-in for the `world_coords` tensor of the reference's video_dict, shape (V,S,S,3) per
    scene or (B,V,S,S,3) after merge_video_dict, backed by raw depth/pose/intrinsics."""

    def __init__(self, depth, poses, intrinsics, crop_size=384, strategy="center_crop", clamp=None, dtype=torch.float32,
                 batched=False, resampled=False, src_hw=None):
        self.depth, self.poses, self.intrinsics = depth, poses, intrinsics
        self.crop_size, self.strategy = crop_size, strategy
        # 'norm' clamp (llava/video_utils.py:232-234): (lo, hi) of ONE scene, or -- batched -- a list with one entry
        # (or None) per scene: the reference clamps every scene to its own mesh AABB before stacking
        self.clamp = clamp
        self.dtype = dtype
        self.batched = batched  # depth is (B,V,H,W) instead of (V,H,W)
        # resampled: depth holds the crop_size x crop_size samples the loader already gathered (scene_pack.py) instead of
        # full frames; src_hw is then the size of the frames they were taken from (it fixes the index maps)
        self.resampled = bool(resampled)
        self.src_hw = tuple(src_hw) if src_hw is not None else tuple(depth.shape[-2:])
        if self.resampled and src_hw is None:
            raise ValueError("resampled depth needs src_hw, the size of the frames it was taken from")
        if batched and clamp is not None and not isinstance(clamp, list):
            raise TypeError("a batched handle takes a list of per-scene clamps")
        self._full = None

    # ---- tensor-like surface used by the reference ------------------------------------------
    @property
    def batch_size(self):
        return self.depth.shape[0] if self.batched else 1

    @property
    def shape(self):
        lead = tuple(self.depth.shape[:-2])
        return torch.Size(lead + (self.crop_size, self.crop_size, 3))

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    @property
    def device(self):
        return self.depth.device

    @property
    def ndim(self):
        return len(self.shape)

    def __len__(self):
        return self.shape[0]

    def _like(self, depth=None, poses=None, intrinsics=None, dtype=None, batched=None, clamp="same"):
        out = LazyWorldCoords(self.depth if depth is None else depth, self.poses if poses is None else poses,
                              self.intrinsics if intrinsics is None else intrinsics, self.crop_size, self.strategy,
                              self.clamp if isinstance(clamp, str) else clamp,
                              self.dtype if dtype is None else dtype, self.batched if batched is None else batched,
                              resampled=self.resampled, src_hw=self.src_hw)
        for name in ("_ready_event", "_loader_slot"):  # set by SceneBatchLoader(pinned_cache=True) on device-resident batches
            if name in self.__dict__:  # (not hasattr: unknown public names fall through to the materialised tensor)
                out.__dict__[name] = self.__dict__[name]
        return out

    def _scene(self, b: int):
        """Scene b of a batched handle, with its own clamp."""
        return self._like(self.depth[b], self.poses[b], self.intrinsics[b], batched=False,
                          clamp=None if self.clamp is None else self.clamp[b])

    def to(self, *args, **kwargs):
        device, dtype = kwargs.get("device"), kwargs.get("dtype")
        for a in args:
            if isinstance(a, torch.dtype):
                dtype = a
            elif isinstance(a, (str, torch.device, int)):
                device = a
            elif isinstance(a, torch.Tensor):
                device, dtype = a.device, a.dtype
        out = self
        if device is not None:
            ev = self.__dict__.get("_ready_event")
            if ev is not None and self.depth.is_cuda:
                torch.cuda.current_stream(self.depth.device).wait_event(ev)  # filled by a loader's copy stream
            nb = bool(kwargs.get("non_blocking", False))
            out = out._like(self.depth.to(device, non_blocking=nb), self.poses.to(device, non_blocking=nb),
                            self.intrinsics.to(device, non_blocking=nb))
        if dtype is not None:
            out = out._like(dtype=dtype)
        return out

    def cuda(self, device=None, non_blocking=False):
        return self.to(device=torch.device("cuda", torch.cuda.current_device() if device is None else device), non_blocking=non_blocking)

    def half(self):
        return self._like(dtype=torch.float16)

    def bfloat16(self):
        return self._like(dtype=torch.bfloat16)

    def float(self):
        return self._like(dtype=torch.float32)

    def detach(self):
        return self

    def pin_memory(self):
        return self._like(self.depth.pin_memory(), self.poses.pin_memory(), self.intrinsics.pin_memory())

    def __iter__(self):
        if not self.batched:
            raise TypeError("iterate a batched LazyWorldCoords (B,V,S,S,3)")
        for b in range(self.depth.shape[0]):
            yield self._scene(b)

    def __getitem__(self, idx):
        if self.batched and isinstance(idx, int):
            return self._scene(idx)
        # world_coords[:, :n, :n, :] of a scene (llava_arch.py:367, 372): the start of the object-proposal chain, deferred so
        # that the unchanged loop ends in one mask kernel straight from depth (lazy_objects.py)
        if (not self.batched and self._full is None and isinstance(idx, tuple) and len(idx) == 4 and all(isinstance(i, slice) for i in idx)
                and idx[0] == slice(None) and idx[3] == slice(None) and idx[1] == idx[2] and idx[1].start is None and idx[1].step is None
                and isinstance(idx[1].stop, int) and 0 < idx[1].stop <= self.crop_size):
            from .lazy_objects import PatchGrid

            return PatchGrid(self, idx[1].stop)
        return self.materialize()[idx]

    def view(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
            shape = tuple(shape[0])
        if len(shape) == 3 and shape[1] == -1 and shape[2] == 3 and shape[0] == self.batch_size:
            return _LazyFlat(self)
        return self.materialize().view(*shape)

    def __getattr__(self, name):
        # anything not modelled above falls back to the real tensor
        if name.startswith("_") or name in ("depth", "poses", "intrinsics"):
            raise AttributeError(name)
        return getattr(self.materialize(), name)

    # ---- evaluation ------------------------------------------------------------------------------
    def _maps(self, dev):
        H, W = self.src_hw
        return geometry.resample_maps(H, W, self.crop_size, self.strategy, dev)

    def _full_depth(self, dev):
        """(N,H,W) frames on `dev`; resampled depth is scattered back to the pixels it was taken from (everything else
        stays zero and is never read through the index maps)."""
        d = self.depth.reshape((-1,) + tuple(self.depth.shape[-2:])).to(dev)
        if not self.resampled:
            return d
        H, W = self.src_hw
        rm, cm = self._maps(dev)
        full = torch.zeros((d.shape[0], H, W), dtype=d.dtype, device=dev)
        lib = cabi.load()
        with torch.cuda.device(dev):
            cabi.check(lib.v3d_scatter_resampled_depth(d.contiguous().data_ptr(), full.data_ptr(), cabi.depth_code(d), d.shape[0], H, W,
                                                       rm.data_ptr(), cm.data_ptr(), d.shape[1], d.shape[2], cabi.stream_ptr(dev)),
                       "v3d_scatter_resampled_depth")
        return full

    def materialize(self) -> torch.Tensor:
        """The tensor the reference would hold: (…,S,S,3) in self.dtype (llava/video_utils.py:296-321)."""
        if self._full is None:
            if self.batched and self.clamp is not None:
                self._full = torch.stack([self._scene(b).materialize() for b in range(self.depth.shape[0])])
                return self._full
            dev = _device_for(self.depth)
            rm, cm = self._maps(dev)
            if self.resampled and tuple(self.depth.shape[-2:]) != (self.crop_size, self.crop_size):
                raise ValueError(f"cannot materialise {tuple(self.shape)} coordinates from {tuple(self.depth.shape[-2:])} resampled depth")
            full = unproject(self.intrinsics.reshape(-1, 4, 4), self.poses.reshape(-1, 4, 4), self._full_depth(dev), row_map=rm, col_map=cm,
                             clamp=self.clamp)
            self._full = full.reshape(tuple(self.shape)).to(self.dtype)
        return self._full

    def pooled(self, patch_size: int, mode: str, trim: int = 6):
        """Pooled coordinates (V,G,G[,n],3) straight from depth, deferred (lazy.LazyPooled): evaluated by the fused kernel
        when -- and with whatever -- the caller finally consumes them (discretisation, PE, token add)."""
        from .lazy import LazyPooled

        if self.batched:
            raise TypeError("pool one scene at a time (the reference iterates over the batch, llava_arch.py:396)")
        return LazyPooled(self, patch_size, mode, trim, cabi.load().v3d_pool_points(_POOL_CODES[mode]))


_POOL_CODES = {"avg": cabi.POOL_AVG, "minmax": cabi.POOL_MINMAX, "sample9": cabi.POOL_SAMPLE9, "sample5": cabi.POOL_SAMPLE5,
               "sample1": cabi.POOL_SAMPLE1}


class VideoProcessor:
    """llava/video_utils.py:71-358 with the same constructor and methods.  `lazy` (extra, keyword
    only) selects what `world_coords` is: a LazyWorldCoords handle (default) or a real tensor."""

    def __init__(self, video_folder="data", annotation_dir="data/embodiedscan/", voxel_size=None, min_xyz_range=None,
                 max_xyz_range=None, frame_sampling_strategy="uniform", val_box_type="pred", *, lazy=True, metadata_dir=None,
                 pack_dir=None, mc_seeds=None):
        self.video_folder = video_folder
        self.voxel_size = voxel_size
        self.min_xyz_range = torch.tensor(min_xyz_range) if min_xyz_range is not None else None
        self.max_xyz_range = torch.tensor(max_xyz_range) if max_xyz_range is not None else None
        self.frame_sampling_strategy = frame_sampling_strategy
        self.lazy = lazy
        # extras (keyword only): packed per-scene binaries (scene_pack.py) replace the PNG/txt reads when present;
        # mc_seeds {video_id: int} fixes the tie-break seed of the online MC sampler (default crc32(video_id))
        self.pack_dir = pack_dir
        self.mc_seeds = dict(mc_seeds or {})
        self._packs = {}
        self._mc_sampler = None
        self.scene = {}
        print("============frame sampling strategy: {}=============".format(self.frame_sampling_strategy))
        # the reference hard-codes data/metadata relative to the working directory (:102,109,116)
        meta = metadata_dir if metadata_dir is not None else os.path.join("data", "metadata")

        for split in ["train", "val", "test"]:
            with open(os.path.join(annotation_dir, f"embodiedscan_infos_{split}.pkl"), "rb") as f:
                for item in pickle.load(f)["data_list"]:
                    if item["sample_idx"].startswith("scannet"):
                        self.scene[item["sample_idx"]] = item

        self.scan2obj = {}
        for split in ["train", "val"]:
            box_type = "gt" if split == "train" else val_box_type
            with open(os.path.join(meta, f"scannet_{split}_{box_type}_box.json")) as f:
                self.scan2obj.update(json.load(f))

        if "mc" in self.frame_sampling_strategy:
            # "online" in the strategy (e.g. mc-online-ratio95): the order is computed here by the CUDA sampler
            # instead of being read from the offline script's JSON
            self.mc_sampling_files = {}
            if "online" not in self.frame_sampling_strategy:
                with open(os.path.join(meta, "scannet_select_frames.json")) as f:
                    self.mc_sampling_files = {dd["video_id"]: dd for dd in json.load(f)}
            with open(os.path.join(meta, "pcd_discrete_0.1.pkl"), "rb") as f:
                pc_data = pickle.load(f)
            self.pc_data = pc_data
            self.pc_min, self.pc_max = {}, {}
            for scene_id, pts in pc_data.items():
                arr = np.asarray(pts, dtype=np.int64).reshape(-1, 3)
                # the reference starts its running min/max at +-1000 (:121-122)
                mn = np.minimum(arr.min(0), 1000) if len(arr) else np.array([1000] * 3)
                mx = np.maximum(arr.max(0), -1000) if len(arr) else np.array([-1000] * 3)
                self.pc_min[scene_id] = torch.Tensor(mn.tolist()) / 10
                self.pc_max[scene_id] = torch.Tensor(mx.tolist()) / 10

    # ---- frame selection (host, integer work) ------------------------------------------------
    def sample_frame_files_mc(self, video_id: str, frames_upbound: int = 32, do_shift=False):
        """:131-159 -- prefix of the offline MC order, cut at the coverage ratio, sorted by frame id."""
        if video_id not in self.mc_sampling_files and "online" in self.frame_sampling_strategy:
            self.mc_sampling_files[video_id] = self.mc_sample_online(video_id)
        mc = self.mc_sampling_files[video_id]
        frame_files = mc["frame_files"][:frames_upbound]
        voxel_nums = mc["voxel_nums"][:frames_upbound]
        ratio = 1.0
        if "ratio95" in self.frame_sampling_strategy:
            ratio = 0.95
        elif "ratio90" in self.frame_sampling_strategy:
            ratio = 0.9
        if ratio != 1.0:
            need = mc["num_all_voxels"] * ratio
            kept, covered = [], 0
            for frame_file, n in zip(frame_files, voxel_nums):
                kept.append(frame_file)
                covered += n
                if covered >= need:
                    break
            frame_files = kept
        frame_files.sort(key=lambda file: int(file.split("/")[-1].split(".")[0]))
        return frame_files

    def sample_frame_files(self, video_id: str, force_sample: bool = False, frames_upbound: int = 0):
        """:162-194 -- uniform np.linspace over the scene's frames (10 unless force_sample)."""
        frame_files = [os.path.join(self.video_folder, img["img_path"]) for img in self.scene[video_id]["images"]]
        n = frames_upbound if force_sample else 10
        picks = np.linspace(0, len(frame_files) - 1, n, dtype=int)
        return [frame_files[i] for i in picks]

    def mc_sample_online(self, video_id: str, max_pick: int = 32):
        """scripts/3d/preprocessing/max_coverage_sampling.py:24-104 for one scene, at run time: candidates =
        every second frame (all if that leaves < 32), voxelise + greedy on the device, the record of :98-104.
        The reference caps the order at 32 picks (:93); so does the kernel's record."""
        from . import mc_sampling

        if torch.utils.data.get_worker_info() is not None and not torch.cuda.is_initialized():
            # CUDA cannot be initialised in a forked DataLoader worker (INTEGRATION.md "Process model")
            raise RuntimeError(
                f"frame_sampling_strategy={self.frame_sampling_strategy!r} computes the max-coverage order on the GPU, which a forked "
                "DataLoader worker cannot do: call VideoProcessor.precompute_mc_order() in the parent process before the loader starts "
                "(the launcher does when V3D_MC_PRECOMPUTE=1), or run with dataloader_num_workers=0 / multiprocessing_context='spawn'")
        info = self.scene[video_id]
        files = [os.path.join(self.video_folder, img["img_path"]) for img in info["images"]][::2]
        if len(files) < 32:
            files = [os.path.join(self.video_folder, img["img_path"]) for img in info["images"]]
        dev = _device_for()
        if self._mc_sampler is None:
            self._mc_sampler = mc_sampling.MCSampler(self.voxel_size if self.voxel_size is not None else 0.1, device=dev)
        depth, poses, intr = self.load_frames(video_id, files)
        scene_id = video_id.split("/")[-1]
        pc = torch.tensor(np.asarray(self.pc_data[scene_id], dtype=np.int32).reshape(-1, 3), device=dev)
        import zlib

        seed = self.mc_seeds.get(video_id, zlib.crc32(video_id.encode()))
        res = self._mc_sampler.sample([{"depth": depth.to(dev), "poses": poses.to(dev), "intrinsics": intr.to(dev), "pc_voxels": pc,
                                        "seed": seed}])[0]
        return {"video_id": video_id, "frame_files": [files[i] for i in res["order"][:max_pick]],
                "voxel_nums": res["gains"][:max_pick], "num_all_voxels": res["n_all"], "num_select_voxels": res["n_sel"]}

    def precompute_mc_order(self, video_ids=None) -> int:
        """Run the online MC sampler for `video_ids` (default: every scene with a mesh universe) NOW, in this process, and
        keep the records: forked DataLoader workers inherit them and never need the GPU.  Returns the number of scenes."""
        if "online" not in self.frame_sampling_strategy:
            return 0
        ids = list(video_ids) if video_ids is not None else [v for v in self.scene if v.split("/")[-1] in self.pc_data]
        for video_id in ids:
            if video_id not in self.mc_sampling_files:
                self.mc_sampling_files[video_id] = self.mc_sample_online(video_id)
        return len(ids)

    # ---- geometry --------------------------------------------------------------------------------
    def _pack(self, video_id: str):
        """The scene's packed binary (scene_pack.py) if pack_dir holds one, else None."""
        if self.pack_dir is None:
            return None
        if video_id not in self._packs:
            from . import scene_pack

            path = scene_pack.pack_path(self.pack_dir, video_id)
            pack = scene_pack.ScenePack(path) if os.path.isfile(path) else None
            # loader threads race here; the FIRST mapping must stay the cached one -- a replaced ScenePack would be unmapped
            # while another thread still copies from its pages (v3d_host_gather takes raw addresses)
            self._packs.setdefault(video_id, pack)
        return self._packs[video_id]

    def load_frames(self, video_id: str, frame_files):
        """Raw inputs of :196-230: uint16 depth stack, float32(axis_align @ pose), float32 intrinsics -- from the
        scene's packed binary when there is one (one contiguous read per frame), else from the PNG/txt files."""
        pack = self._pack(video_id)
        if pack is not None:
            from . import scene_pack

            return pack.load([scene_pack.frame_number(f) for f in frame_files])
        return self.load_frames_from_files(video_id, frame_files)

    def world_coords(self, video_id: str, frame_files, crop_size: int = 384, strategy: str = "center_crop", do_normalize=None):
        """The `world_coords` entry of preprocess() (:291-321) for a given frame list, without touching the RGB frames:
        a LazyWorldCoords handle (lazy=True) or the (V,crop,crop,3) tensor.  With a packed scene that stores the depth
        already resampled for (crop_size, strategy), a lazy handle carries those samples (48 % of the bytes) and the
        fused kernel reads them in place."""
        if do_normalize is None:
            do_normalize = "norm" in self.frame_sampling_strategy
        clamp = self._clamp_for(video_id, do_normalize)
        pack = self._pack(video_id)
        if self.lazy and pack is not None and pack.resampled_for == (crop_size, strategy):
            from . import scene_pack

            depth, poses, intr = pack.load_resampled([scene_pack.frame_number(f) for f in frame_files])
            return LazyWorldCoords(depth, poses, intr, crop_size=crop_size, strategy=strategy, clamp=clamp, resampled=True,
                                   src_hw=(pack.H, pack.W))
        depth, poses, intr = self.load_frames(video_id, frame_files)
        handle = LazyWorldCoords(depth, poses, intr, crop_size=crop_size, strategy=strategy, clamp=clamp)
        return handle if self.lazy else handle.materialize()

    def load_frames_from_files(self, video_id: str, frame_files):
        """:212-230 as written: Image.open of the depth PNG, np.loadtxt of the pose, axis_align(f64) @ pose."""
        meta_info = self.scene[video_id]
        axis_align = np.array(meta_info["axis_align_matrix"], dtype=np.float64)
        intr = np.array(meta_info["depth_cam2img"], dtype=np.float64)
        depths, poses = [], []
        for frame_path in frame_files:
            with Image.open(frame_path.replace(".jpg", ".png")) as im:
                d = np.array(im)
            if d.dtype != np.uint16:
                d = d.astype(np.int32)
            depths.append(d)
            poses.append(axis_align @ np.loadtxt(frame_path.replace("jpg", "txt")))
        depth = np.stack(depths)
        depth_t = torch.from_numpy(depth.view(np.int16)) if depth.dtype == np.uint16 else torch.from_numpy(depth)
        poses_t = torch.from_numpy(np.stack(poses)).float()
        intr_t = torch.from_numpy(intr).unsqueeze(0).repeat(len(frame_files), 1, 1).float()
        return depth_t, poses_t, intr_t

    def _clamp_for(self, video_id: str, do_normalize: bool):
        if not do_normalize:
            return None
        scene_id = video_id.split("/")[-1]
        return (self.pc_min[scene_id].tolist(), self.pc_max[scene_id].tolist())

    def calculate_world_coords(self, video_id: str, frame_files, do_normalize=False):
        """:196-238 -> {"world_coords": (V,H,W,3) float32} on the CUDA device (always eager: the MC
        preprocessing script consumes full-resolution coordinates, max_coverage_sampling.py:34-37)."""
        depth, poses, intr = self.load_frames(video_id, frame_files)
        return {"world_coords": unproject(intr, poses, depth, clamp=self._clamp_for(video_id, do_normalize))}

    def preprocess(self, video_id: str, image_processor, force_sample: bool = False, frames_upbound: int = 0,
                   strategy: str = "center_crop"):
        """:242-326 -> {"images", "world_coords", "video_size", "boundry", "objects"}."""
        if "mc" in self.frame_sampling_strategy:
            frame_files = self.sample_frame_files_mc(video_id, frames_upbound=frames_upbound,
                                                     do_shift=("shift" in self.frame_sampling_strategy))
        else:
            frame_files = self.sample_frame_files(video_id, force_sample=force_sample, frames_upbound=frames_upbound)
        pack = self._pack(video_id)
        if pack is not None:
            H, W = pack.H, pack.W
        else:
            with Image.open(frame_files[0].replace(".jpg", ".png")) as im:
                W, H = im.size

        images = []
        for frame_file in frame_files:
            with Image.open(frame_file) as img:
                images.append(img.convert("RGB"))
        crop_size = image_processor.crop_size["width"]
        if strategy == "resize":
            images = [frame.resize((crop_size, crop_size)) for frame in images]
        elif strategy == "center_crop":
            new_height, new_width = crop_size, int(W * (crop_size / H))
            left, top = (new_width - crop_size) // 2, (new_height - crop_size) // 2
            images = [frame.resize((new_width, new_height)).crop((left, top, left + crop_size, top + crop_size)) for frame in images]

        clamp = self._clamp_for(video_id, "norm" in self.frame_sampling_strategy)
        world_coords = self.world_coords(video_id, frame_files, crop_size, strategy)
        if self.lazy:
            # `boundry` (:273-277) is dropped by merge_video_dict: evaluated (from the full-resolution frames) only if read
            boundry = _LazyBoundary(lambda: self.load_frames(video_id, frame_files), clamp)
        else:
            depth, poses, intr = self.load_frames(video_id, frame_files)
            boundry = boundary(intr, poses, depth, clamp)
        return {
            "images": images,
            "world_coords": world_coords,
            "video_size": len(images),
            "boundry": boundry,
            "objects": torch.tensor(self.scan2obj[video_id]),
        }

    def process_3d_video(self, video_id: str, image_processor, force_sample: bool = False, frames_upbound: int = 0,
                         strategy: str = "center_crop"):
        """:329-345 -- preprocess + SigLIP image preprocessing of the RGB frames."""
        video_dict = self.preprocess(video_id, image_processor, force_sample, frames_upbound, strategy)
        video_dict["images"] = image_processor.preprocess(video_dict["images"], return_tensors="pt")["pixel_values"]
        return video_dict

    def discrete_point(self, xyz):
        """:348-358 -- host twin of discrete_coords on a list of points; returns nested int lists.
        A handful of points per call (box centres, the origin): plain torch on the host, as upstream."""
        xyz = torch.tensor(xyz)
        if self.min_xyz_range is not None:
            xyz = torch.maximum(xyz, self.min_xyz_range.to(xyz.device))
        if self.max_xyz_range is not None:
            xyz = torch.minimum(xyz, self.max_xyz_range.to(xyz.device))
        if self.min_xyz_range is not None:
            xyz = xyz - self.min_xyz_range.to(xyz.device)
        xyz = xyz / self.voxel_size
        return xyz.round().int().tolist()


class _LazyBoundary:
    """`boundry` is dropped by merge_video_dict and unused downstream; compute on demand."""

    def __init__(self, load, clamp):
        self._load, self._clamp, self._val = load, clamp, None

    def tensor(self) -> torch.Tensor:
        if self._val is None:
            d, p, k = self._load()
            self._val = boundary(k, p, d, self._clamp).cpu()
        return self._val

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.tensor(), name)

    def __getitem__(self, i):
        return self.tensor()[i]


def merge_video_dict(video_dict_list):
    """llava/video_utils.py:361-373: stack world_coords/images/objects over the batch, collect the
    non-None box_input entries, drop every other key.  Lazy handles are merged into one batched handle."""
    new_video_dict = {"box_input": []}
    for k in video_dict_list[0]:
        if k in ["world_coords", "images", "objects"]:
            vals = [vd[k] for vd in video_dict_list]
            if k == "world_coords" and isinstance(vals[0], LazyWorldCoords):
                h0 = vals[0]
                for v in vals[1:]:
                    same = (v.crop_size, v.strategy, v.dtype, v.resampled, v.src_hw) == (h0.crop_size, h0.strategy, h0.dtype, h0.resampled, h0.src_hw)
                    if not isinstance(v, LazyWorldCoords) or not same:
                        raise ValueError("cannot stack world_coords handles with different crop size / strategy / dtype / layout")
                clamps = [v.clamp for v in vals]
                new_video_dict[k] = LazyWorldCoords(torch.stack([v.depth for v in vals]), torch.stack([v.poses for v in vals]),
                                                    torch.stack([v.intrinsics for v in vals]), h0.crop_size, h0.strategy,
                                                    None if all(c is None for c in clamps) else clamps, h0.dtype, batched=True,
                                                    resampled=h0.resampled, src_hw=h0.src_hw)
            else:
                new_video_dict[k] = torch.stack(vals)
        elif k in ["box_input"]:
            for vd in video_dict_list:
                if vd[k] is not None:
                    new_video_dict["box_input"].append(vd[k])
    new_video_dict["box_input"] = torch.Tensor(new_video_dict["box_input"])
    return new_video_dict


class SceneBatchLoader:
    """Background loader of `world_coords` batches for a list of samples: what a DataLoader with
    pin_memory=True and a collate of merge_video_dict produces, without the per-sample tensors and the stack copy --
    every scene's frames go from its packed binary straight into a ring of pinned batch buffers, one native
    multi-threaded gather per batch (v3d_host_gather), `prefetch` batches ahead of the consumer.

        loader = SceneBatchLoader(vp, samples, batch_scenes=16)        # samples: [(video_id, frame_files), ...]
        for handle in loader:                                          # batched LazyWorldCoords over pinned memory
            dev = handle.to("cuda", non_blocking=True)                 # one DMA per tensor
            for scene in dev: ...                                      # llava_arch.py:396 iterates the batch the same way

    A yielded handle stays valid until `prefetch` further batches have been requested.  Every sample needs a packed scene
    (VideoProcessor(pack_dir=...)) and the same number of frames; with packs that store the resampled depth for
    (crop_size, strategy) the handles carry those samples (48 % of the bytes of full ScanNet frames)."""

    def __init__(self, vp: "VideoProcessor", samples, batch_scenes: int, crop_size: int = 384, strategy: str = "center_crop",
                 threads: int = 8, prefetch: int = 2, loop: bool = False, pin: bool = True, pinned_cache: bool = False, device=None):
        from concurrent.futures import ThreadPoolExecutor

        from . import scene_pack

        if not samples:
            raise ValueError("no samples")
        self.vp, self.samples, self.B = vp, list(samples), int(batch_scenes)
        self.crop_size, self.strategy, self.loop, self.prefetch = crop_size, strategy, loop, max(1, int(prefetch))
        self.V = len(self.samples[0][1])
        if any(len(f) != self.V for _, f in self.samples):
            raise ValueError("every sample of a batch loader needs the same number of frames (the reference stacks them, video_utils.py:367)")
        p0 = vp._pack(self.samples[0][0])
        if p0 is None:
            raise ValueError(f"no packed scene for {self.samples[0][0]} under pack_dir={vp.pack_dir!r}")
        self.resampled = p0.resampled_for == (crop_size, strategy)
        self.src_hw = (p0.H, p0.W)
        shape = (crop_size, crop_size) if self.resampled else (p0.H, p0.W)
        store = torch.int16 if p0.depth_dtype == np.uint16 else torch.int32
        self._frame_numbers = [[scene_pack.frame_number(f) for f in files] for _, files in self.samples]
        # pinned_cache: every scene's stack is copied into pinned memory once (ScenePack.pin) and a batch is gathered by the
        # DMA engines straight into a ring of DEVICE buffers on the loader's own stream -- no CPU copy per batch; the
        # yielded handles are device-resident and carry the event to wait for (LazyWorldCoords.to does)
        self.pinned_cache = bool(pinned_cache)
        self._dev = None
        if self.pinned_cache:
            self._dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
            self._stream = torch.cuda.Stream(device=self._dev)
            self._consumed = [None] * (self.prefetch + 2)
            self._ring = [(torch.empty((self.B, self.V) + shape, dtype=store, device=self._dev),
                           torch.empty((self.B, self.V, 4, 4), dtype=torch.float32, pin_memory=True),
                           torch.empty((self.B, self.V, 4, 4), dtype=torch.float32, pin_memory=True)) for _ in range(self.prefetch + 2)]
        else:
            self._ring = [(torch.empty((self.B, self.V) + shape, dtype=store, pin_memory=pin),
                           torch.empty((self.B, self.V, 4, 4), dtype=torch.float32, pin_memory=pin),
                           torch.empty((self.B, self.V, 4, 4), dtype=torch.float32, pin_memory=pin)) for _ in range(self.prefetch + 1)]
        # batches are assembled by `prefetch` Python threads, each issuing ONE native multi-threaded gather per batch
        self._pool = ThreadPoolExecutor(max_workers=self.prefetch, thread_name_prefix="v3d-load")
        self.threads = max(1, int(threads))
        self._lib = cabi.load()
        self._do_norm = "norm" in vp.frame_sampling_strategy

    def __len__(self):
        return len(self.samples) // self.B

    def _load_batch(self, batch: int, slot: int):
        """Fill ring slot `slot` with batch `batch`: ONE native gather of every frame of every scene (v3d_host_gather, the
        interpreter lock is released for its duration), poses / intrinsics with numpy.  Returns the per-scene clamps."""
        import ctypes as C

        depth, poses, intr = self._ring[slot]
        n = len(self.samples)
        srcs = (C.c_void_p * (self.B * self.V))()
        clamps = []
        frame_bytes = depth[0, 0].numel() * depth.element_size()
        for b in range(self.B):
            idx = (batch * self.B + b) % n
            video_id = self.samples[idx][0]
            pack = self.vp._pack(video_id)
            if pack is None or (pack.H, pack.W) != self.src_hw or (pack.resampled_for == (self.crop_size, self.strategy)) != self.resampled:
                raise ValueError(f"{video_id}: packed scene missing or of a different layout than the first sample's")
            if self.pinned_cache and not pack.pinned:
                pack.pin()  # once per scene
            rows = pack.rows(self._frame_numbers[idx])
            src = pack.resampled if self.resampled else pack.depth
            base = src.ctypes.data
            for o, r in enumerate(rows):
                srcs[b * self.V + o] = base + int(r) * frame_bytes
            poses[b].numpy()[...] = pack.poses[rows]
            intr[b].numpy()[...] = pack.intr[None]
            clamps.append(self.vp._clamp_for(video_id, self._do_norm))
        if self.pinned_cache:
            with torch.cuda.device(self._dev):
                if self._consumed[slot] is not None:
                    self._stream.wait_event(self._consumed[slot])  # the consumer's kernels on the slot's previous batch
                cabi.check(self._lib.v3d_h2d_gather(depth.data_ptr(), srcs, frame_bytes, self.B * self.V, self._stream.cuda_stream), "v3d_h2d_gather")
                with torch.cuda.stream(self._stream):
                    poses_d, intr_d = poses.to(self._dev, non_blocking=True), intr.to(self._dev, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self._stream)
            return clamps, poses_d, intr_d, ev
        cabi.check(self._lib.v3d_host_gather(depth.data_ptr(), srcs, frame_bytes, self.B * self.V, self.threads), "v3d_host_gather")
        return clamps

    def _submit(self, batch: int, slot: int):
        return self._pool.submit(self._load_batch, batch, slot)

    def __iter__(self):
        n_batches = len(self)
        if n_batches == 0:
            return
        pending = []  # (slot, futures) in batch order
        nxt = 0
        while True:
            while len(pending) < self.prefetch and (self.loop or nxt < n_batches):
                slot = nxt % len(self._ring)
                pending.append((slot, self._submit(nxt % n_batches if self.loop else nxt, slot)))
                nxt += 1
            if not pending:
                return
            slot, fut = pending.pop(0)
            res = fut.result()
            depth, poses, intr = self._ring[slot]
            ev = None
            if self.pinned_cache:
                clamps, poses, intr, ev = res
            else:
                clamps = res
            h = LazyWorldCoords(depth, poses, intr, self.crop_size, self.strategy, None if all(c is None for c in clamps) else clamps,
                                batched=True, resampled=self.resampled, src_hw=self.src_hw)
            if ev is not None:
                h._ready_event, h._loader_slot = ev, slot
            yield h

    def mark_consumed(self, handle) -> None:
        """pinned_cache mode: call after the last launch that reads a yielded handle (on the current stream); the loader's
        copies into the handle's ring slot wait for it.  Without it a handle must not be in use any more when
        prefetch + 1 further batches have been requested."""
        slot = handle.__dict__.get("_loader_slot")
        if self.pinned_cache and slot is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self._dev))
            self._consumed[slot] = ev

    def close(self):
        """Cancel queued batches and WAIT for the one being gathered: it writes into the pinned ring, which must outlive it."""
        self._pool.shutdown(wait=True, cancel_futures=True)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

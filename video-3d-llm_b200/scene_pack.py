"""Packed per-scene binary of the geometry inputs.

The reference reads, per sampled frame, one 16-bit PNG (`Image.open` + `np.array`, llava/video_utils.py:215-218)
and one `%f` text pose (`np.loadtxt`, :221-223) and multiplies by the axis-align matrix in float64 (:205,227).
With the device side at microseconds per frame that host work is the end-to-end bound, so a scene is converted
once into one file that holds exactly what `VideoProcessor.load_frames` produces:

    offset 0     header, 4096 bytes, little endian:
                   magic "V3DPACK1", u32 version, u32 n_frames, u32 H, u32 W, u32 depth_code (0 = uint16,
                   1 = int32), u32 reserved, u64 off_ids, u64 off_poses, u64 off_intr, u64 off_depth, u64 file_bytes,
                   (version 2:) u64 off_resampled (0 = none), u32 crop_size, u32 strategy (0 = center_crop, 1 = resize)
    off_ids      int32[n_frames]        frame number (file stem, e.g. 00120.png -> 120)
    off_poses    float32[n_frames][16]  float32(axis_align(f64) @ loadtxt(pose)(f64)), the tensor of :230
    off_intr     float32[16]            depth_cam2img as float32 (:206,228)
    off_depth    depth[n_frames][H][W]  4096-byte aligned; frame i is one contiguous H*W*itemsize block
    off_resampled  depth[n_frames][crop][crop]  optional, 4096-byte aligned: the nearest-neighbour resize + centre crop of
                 :291-308 applied to the raw depth at conversion time (depth[rows][:, cols] with the index maps of geometry.py)
                 -- exactly the samples the model-side front end reads (48 % of a ScanNet frame), so a training sample is one
                 contiguous copy per frame into pinned memory and the fused kernel reads it in place
                 (v3d_frontend_desc.depth_resampled); the full-resolution stack stays for the MC sampler and for boundary()

Reading a set of frames is one contiguous copy per frame from the page cache (np.memmap) into a pinned staging
buffer, then a single async H2D copy.  Bit-identical to the PNG/txt path (tests/test_scene_pack.py).

    python -m video_3d_llm_b200.scene_pack --video_dir data/scannet/posed_images --out_dir data/scannet/packed
"""
from __future__ import annotations

import argparse
import os
import struct

import numpy as np
import torch

MAGIC = b"V3DPACK1"
VERSION = 2
HEADER_BYTES = 4096
_HDR = struct.Struct("<8sIIIIIIQQQQQ")
_HDR2 = struct.Struct("<QII")  # follows _HDR in version 2
_STRATEGIES = ("center_crop", "resize")


def _align(n: int, a: int = 4096) -> int:
    return (n + a - 1) // a * a


def frame_number(path: str) -> int:
    return int(os.path.basename(path).split(".")[0])


def write_pack(path: str, frame_ids, poses32: np.ndarray, intr32: np.ndarray, depth: np.ndarray, resample=None) -> None:
    """poses32 (n,4,4) float32, intr32 (4,4) float32, depth (n,H,W) uint16 or int32.
    resample = (crop_size, strategy) also stores the resampled stack (see the module docstring)."""
    from . import geometry

    n, H, W = depth.shape
    if depth.dtype == np.uint16:
        code = 0
    elif depth.dtype == np.int32:
        code = 1
    else:
        raise ValueError(f"depth must be uint16 or int32, got {depth.dtype}")
    ids = np.asarray(frame_ids, np.int32)
    poses32 = np.ascontiguousarray(poses32, np.float32).reshape(n, 16)
    intr32 = np.ascontiguousarray(intr32, np.float32).reshape(16)
    if ids.shape != (n,):
        raise ValueError("one frame id per depth frame")
    off_ids = HEADER_BYTES
    off_poses = off_ids + _align(4 * n, 64)
    off_intr = off_poses + 64 * n
    off_depth = _align(off_intr + 64)
    off_res, crop, strat, res = 0, 0, 0, None
    if resample is not None:
        crop, strategy = int(resample[0]), resample[1]
        strat = _STRATEGIES.index(strategy)
        rows, cols = geometry.resample_maps_host(H, W, crop, strategy)
        res = np.ascontiguousarray(depth[:, np.asarray(rows)][:, :, np.asarray(cols)])
        off_res = _align(off_depth + depth.nbytes)
    total = off_depth + depth.nbytes if res is None else off_res + res.nbytes
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        hdr = _HDR.pack(MAGIC, VERSION, n, H, W, code, 0, off_ids, off_poses, off_intr, off_depth, total) + _HDR2.pack(off_res, crop, strat)
        f.write(hdr.ljust(HEADER_BYTES, b"\0"))
        f.write(ids.tobytes().ljust(off_poses - off_ids, b"\0"))
        f.write(poses32.tobytes())
        f.write(intr32.tobytes().ljust(off_depth - off_intr, b"\0"))
        f.write(np.ascontiguousarray(depth).tobytes())
        if res is not None:
            f.write(b"\0" * (off_res - off_depth - depth.nbytes))
            f.write(res.tobytes())
    os.replace(tmp, path)


class ScenePack:
    """Read side: memory-mapped views, frames addressed by frame number."""

    def __init__(self, path: str):
        self.path = path
        with open(path, "rb") as f:
            hdr = f.read(_HDR.size + _HDR2.size)
        if len(hdr) < _HDR.size + _HDR2.size:
            raise ValueError(f"{path}: truncated header")
        magic, ver, n, H, W, code, _, off_ids, off_poses, off_intr, off_depth, total = _HDR.unpack(hdr[:_HDR.size])
        if magic != MAGIC or ver not in (1, 2):
            raise ValueError(f"{path}: not a V3DPACK1 file (magic {magic!r}, version {ver})")
        off_res, crop, strat = _HDR2.unpack(hdr[_HDR.size:]) if ver >= 2 else (0, 0, 0)
        if code not in (0, 1):
            raise ValueError(f"{path}: unknown depth code {code}")
        if os.path.getsize(path) != total:
            raise ValueError(f"{path}: size {os.path.getsize(path)} != {total} recorded in the header")
        self.n_frames, self.H, self.W = n, H, W
        self.depth_dtype = np.uint16 if code == 0 else np.int32
        mm = np.memmap(path, dtype=np.uint8, mode="r")
        self.ids = np.frombuffer(mm, np.int32, n, off_ids)
        self.poses = np.frombuffer(mm, np.float32, 16 * n, off_poses).reshape(n, 4, 4)
        self.intr = np.frombuffer(mm, np.float32, 16, off_intr).reshape(4, 4)
        self.depth = np.frombuffer(mm, self.depth_dtype, n * H * W, off_depth).reshape(n, H, W)
        # optional resampled stack: (crop_size, strategy) it was made for, (n, crop, crop) samples
        self.resampled_for = (crop, _STRATEGIES[strat]) if off_res else None
        self.resampled = np.frombuffer(mm, self.depth_dtype, n * crop * crop, off_res).reshape(n, crop, crop) if off_res else None
        self._index = {int(v): i for i, v in enumerate(self.ids)}

    def pin(self) -> None:
        """Copy the stack the loader reads (the resampled one if the pack has it, else the full frames) into pinned host
        memory once: batches are then gathered by the DMA engines straight out of it (v3d_h2d_gather) and the CPU moves no
        depth.  For datasets whose (resampled) depth fits in host memory -- 286 KB per frame for the ScanNet recipe."""
        if getattr(self, "_pinned", None) is not None:
            return
        src = self.resampled if self.resampled is not None else self.depth
        store = torch.int16 if self.depth_dtype == np.uint16 else torch.int32
        t = torch.empty(src.shape, dtype=store, pin_memory=True)
        t.numpy().view(self.depth_dtype)[...] = src
        self._pinned = t
        if self.resampled is not None:
            self.resampled = t.numpy().view(self.depth_dtype)
        else:
            self.depth = t.numpy().view(self.depth_dtype)

    @property
    def pinned(self) -> bool:
        return getattr(self, "_pinned", None) is not None

    def rows(self, frame_numbers) -> np.ndarray:
        try:
            return np.fromiter((self._index[int(k)] for k in frame_numbers), np.int64)
        except KeyError as e:
            raise KeyError(f"{self.path}: frame {e.args[0]} is not in the pack") from None

    def load_resampled(self, frame_numbers, pin: bool = False, out: torch.Tensor | None = None):
        """(depth (V,crop,crop), poses, intrinsics) from the stored resampled stack: one contiguous copy per frame.
        out: optional (V,crop,crop) destination (e.g. a slice of a pinned ring buffer)."""
        if self.resampled is None:
            raise ValueError(f"{self.path}: the pack holds no resampled stack (write it with resample=(crop_size, strategy))")
        rows = self.rows(frame_numbers)
        V, crop = len(rows), self.resampled.shape[1]
        store = torch.int16 if self.depth_dtype == np.uint16 else torch.int32
        depth_t = out if out is not None else torch.empty((V, crop, crop), dtype=store, pin_memory=pin)
        dst = depth_t.numpy().view(self.depth_dtype)
        for o, r in enumerate(rows):
            dst[o] = self.resampled[r]
        poses_t = torch.from_numpy(np.array(self.poses[rows]))
        intr_t = torch.from_numpy(np.array(self.intr)).unsqueeze(0).repeat(V, 1, 1)
        if pin:
            poses_t, intr_t = poses_t.pin_memory(), intr_t.pin_memory()
        return depth_t, poses_t, intr_t

    def load(self, frame_numbers, pin: bool = False, window=None, resample=None):
        """(depth, poses, intrinsics) tensors for the given frame numbers, like VideoProcessor.load_frames:
        depth (V,H,W) int16 storage of uint16 (or int32), poses (V,4,4) f32, intrinsics (V,4,4) f32.
        window = FrontEnd.upload_window(itemsize) = (x_off_bytes, width_bytes, row_period, rows_keep): copy only the
        samples the front end reads, packed as (V, H//period*keep, width) for FrontEnd.upload_packed.
        resample = (rows, cols) index sequences: the nearest-neighbour resize + crop applied here, on the host, to the raw
        depth -- depth[rows][:, cols] per frame, for FrontEnd.scatter_resampled."""
        rows = self.rows(frame_numbers)
        V = len(rows)
        store = torch.int16 if self.depth_dtype == np.uint16 else torch.int32
        if resample is not None:
            rows_i, cols_i = np.asarray(resample[0], np.int64), np.asarray(resample[1], np.int64)
            if rows_i.min() < 0 or rows_i.max() >= self.H or cols_i.min() < 0 or cols_i.max() >= self.W:
                raise ValueError(f"resample maps reach outside {self.H}x{self.W} frames")
            depth_t = torch.empty((V, len(rows_i), len(cols_i)), dtype=store, pin_memory=pin)
            dst = depth_t.numpy().view(self.depth_dtype)
            for o, r in enumerate(rows):
                dst[o] = self.depth[r][rows_i][:, cols_i]
        elif window is not None:
            es = np.dtype(self.depth_dtype).itemsize
            x_off, width, period, keep = window
            if self.H % period or x_off % es or width % es or x_off + width > self.W * es:
                raise ValueError(f"window {window} does not fit {self.H}x{self.W} frames")
            depth_t = torch.empty((V, self.H // period * keep, width // es), dtype=store, pin_memory=pin)
            dst = depth_t.numpy().view(self.depth_dtype).reshape(V, self.H // period, keep, width // es)
            for o, r in enumerate(rows):
                dst[o] = self.depth[r].reshape(self.H // period, period, self.W)[:, :keep, x_off // es:(x_off + width) // es]
        else:
            depth_t = torch.empty((V, self.H, self.W), dtype=store, pin_memory=pin)
            dst = depth_t.numpy().view(self.depth_dtype)
            for o, r in enumerate(rows):  # one contiguous H*W block per frame
                dst[o] = self.depth[r]
        poses_t = torch.from_numpy(np.array(self.poses[rows]))
        intr_t = torch.from_numpy(np.array(self.intr)).unsqueeze(0).repeat(V, 1, 1)
        if pin:
            poses_t, intr_t = poses_t.pin_memory(), intr_t.pin_memory()
        return depth_t, poses_t, intr_t


def pack_path(pack_dir: str, video_id: str) -> str:
    return os.path.join(pack_dir, video_id.split("/")[-1] + ".v3dpack")


def pack_scene(vp, video_id: str, out_path: str, resample=None) -> str:
    """Convert one scene of a VideoProcessor's dataset (all frames of its info entry) from PNG/txt.
    resample = (crop_size, strategy): also store the resampled stack for the model-side front end."""
    files = [os.path.join(vp.video_folder, img["img_path"]) for img in vp.scene[video_id]["images"]]
    depth_t, poses_t, intr_t = vp.load_frames_from_files(video_id, files)
    depth = depth_t.numpy()
    depth = depth.view(np.uint16) if depth.dtype == np.int16 else depth
    write_pack(out_path, [frame_number(f) for f in files], poses_t.numpy(), intr_t[0].numpy(), depth, resample=resample)
    return out_path


def main(args) -> None:
    from .video_utils import VideoProcessor

    vp = VideoProcessor(video_folder=args.video_folder, annotation_dir=args.annotation_dir, voxel_size=0.1, lazy=True,
                        metadata_dir=args.metadata_dir)
    os.makedirs(args.out_dir, exist_ok=True)
    for scene_id in sorted(os.listdir(args.video_dir)):
        video_id = f"scannet/{scene_id}"
        if video_id not in vp.scene:
            continue
        out = pack_scene(vp, video_id, pack_path(args.out_dir, video_id),
                         resample=(args.resample, args.strategy) if args.resample else None)
        print(f"[{scene_id}] {os.path.getsize(out) / 1e6:.1f} MB -> {out}")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--video_dir", default="data/scannet/posed_images")
    ap.add_argument("--video_folder", default="data")
    ap.add_argument("--annotation_dir", default="data/embodiedscan/")
    ap.add_argument("--metadata_dir", default=None)
    ap.add_argument("--out_dir", default="data/scannet/packed")
    ap.add_argument("--resample", type=int, default=384,
                    help="also store the depth resampled to this crop size (the image processor's crop, 384 for SigLIP); 0 = full frames only")
    ap.add_argument("--strategy", default="center_crop", choices=list(_STRATEGIES))
    main(ap.parse_args())

"""Deterministic synthetic ScanNet-shaped scenes.

No dataset or checkpoint is reachable offline, so tests and bench.py run on scenes made
here: an axis-aligned room with a few interior boxes, a camera on a closed loop, depth as
z-depth in uint16 millimetres with a fraction of invalid (zero) pixels, ScanNet-like
intrinsics and camera-to-world poses already composed with the axis-align matrix
(llava/video_utils.py:205,227), plus the voxelised mesh universe the MC sampler needs
(scripts/3d/preprocessing/convert_pcd_to_voxel.py:69-82).

The per-pixel ray casting runs in torch on the device asked for (CPU in tests, the GPU in
bench.py); it only produces inputs and is not part of any measured path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

FX = FY = 577.87
CX, CY = 319.5, 239.5


@dataclass
class Scene:
    seed: int
    depth: torch.Tensor  # (V,H,W) int16 storage of uint16 millimetres (torch.uint16 is partial)
    poses: torch.Tensor  # (V,4,4) float32, axis_align @ cam2world
    intrinsics: torch.Tensor  # (V,4,4) float32
    room: np.ndarray  # (2,3) float64 lo/hi of the room AABB
    boxes: np.ndarray  # (K,2,3) float64 interior boxes
    axis_align: np.ndarray  # (4,4) float64
    sensor_poses: np.ndarray  # (V,4,4) float64, inv(axis_align) @ (aligned cam2world)

    def depth_u16(self) -> np.ndarray:
        return self.depth.cpu().numpy().view(np.uint16)


def _rot_z(a: float) -> np.ndarray:
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def scene_layout(seed: int, max_boxes: int = 8):
    """Room + interior boxes + axis-align matrix for scene `seed` (rng seed 1000+seed)."""
    rng = np.random.default_rng(1000 + seed)
    sx, sy, sz = rng.uniform(3, 8), rng.uniform(3, 8), rng.uniform(2.4, 3.2)
    room = np.array([[-sx / 2, -sy / 2, 0.0], [sx / 2, sy / 2, sz]])
    k = int(rng.integers(0, max_boxes + 1))
    boxes = []
    for _ in range(k):
        w, d, h = rng.uniform(0.3, 1.2), rng.uniform(0.3, 1.2), rng.uniform(0.3, 1.6)
        # keep furniture away from the camera loop (radius 0.25*min(sx,sy) around the origin)
        r_cam = 0.25 * min(sx, sy)
        for _try in range(20):
            cx = rng.uniform(room[0, 0] + w / 2, room[1, 0] - w / 2)
            cy = rng.uniform(room[0, 1] + d / 2, room[1, 1] - d / 2)
            if math.hypot(abs(cx) - w / 2, abs(cy) - d / 2) > r_cam + 0.3 or min(abs(cx) - w / 2, abs(cy) - d / 2) > r_cam + 0.2:
                break
        boxes.append([[cx - w / 2, cy - d / 2, 0.0], [cx + w / 2, cy + d / 2, h]])
    boxes = np.array(boxes, np.float64).reshape(-1, 2, 3)
    axis_align = np.eye(4)
    axis_align[:3, :3] = _rot_z(rng.uniform(0, 2 * math.pi))
    axis_align[:3, 3] = rng.uniform(-2, 2, 3)
    return rng, room, boxes, axis_align


def camera_loop(rng, room, n_total: int) -> np.ndarray:
    """(n_total,4,4) float64 aligned camera-to-world poses, x right / y down / z forward."""
    sx, sy = room[1, 0] - room[0, 0], room[1, 1] - room[0, 1]
    radius = 0.25 * min(sx, sy)
    out = np.zeros((n_total, 4, 4))
    for i in range(n_total):
        phase = 2 * math.pi * i / max(n_total, 1)
        pos = np.array([radius * math.cos(phase), radius * math.sin(phase), 1.4 + rng.uniform(-0.1, 0.1)])
        yaw = 2 * phase + rng.normal(0, 0.05)
        pitch = -0.3 + rng.normal(0, 0.05)
        f = np.array([math.cos(yaw) * math.cos(pitch), math.sin(yaw) * math.cos(pitch), math.sin(pitch)])
        r = np.cross(f, np.array([0.0, 0.0, 1.0]))
        r /= np.linalg.norm(r)
        d = np.cross(f, r)
        out[i, :3, 0], out[i, :3, 1], out[i, :3, 2], out[i, :3, 3] = r, d, f, pos
        out[i, 3, 3] = 1.0
    return out


def _pixel_hash(seed: int, V: int, H: int, W: int, device) -> torch.Tensor:
    """Cheap deterministic per-pixel hash in [0, 2^16): same bits on every device."""
    v = torch.arange(V, device=device, dtype=torch.int64).view(V, 1, 1)
    r = torch.arange(H, device=device, dtype=torch.int64).view(1, H, 1)
    c = torch.arange(W, device=device, dtype=torch.int64).view(1, 1, W)
    x = (v * 73856093 + r * 19349663 + c * 83492791 + (seed + 1) * 2654435761) & 0xFFFFFFFF
    x = ((x ^ (x >> 15)) * 2246822519) & 0xFFFFFFFF
    x = ((x ^ (x >> 13)) * 3266489917) & 0xFFFFFFFF
    return (x ^ (x >> 16)) & 0xFFFF


def render_depth(c2w: np.ndarray, room: np.ndarray, boxes: np.ndarray, H: int, W: int, device,
                 fx=FX, fy=FY, cx=CX, cy=CY) -> torch.Tensor:
    """z-depth (metres, float32) of the first surface along each pixel ray. c2w (V,4,4) aligned."""
    V = c2w.shape[0]
    R = torch.tensor(c2w[:, :3, :3], dtype=torch.float32, device=device)
    o = torch.tensor(c2w[:, :3, 3], dtype=torch.float32, device=device)
    xc = ((torch.arange(W, device=device, dtype=torch.float32) - cx) / fx).view(1, 1, W)
    yc = ((torch.arange(H, device=device, dtype=torch.float32) - cy) / fy).view(1, H, 1)
    d = [R[:, i, 0].view(V, 1, 1) * xc + R[:, i, 1].view(V, 1, 1) * yc + R[:, i, 2].view(V, 1, 1) for i in range(3)]
    big = torch.full((V, H, W), 1e9, dtype=torch.float32, device=device)
    depth = big.clone()
    for i in range(3):  # camera is inside the room: exit distance per axis
        lo, hi = float(room[0, i]), float(room[1, i])
        oi = o[:, i].view(V, 1, 1)
        t = torch.where(d[i] > 0, (hi - oi) / d[i], torch.where(d[i] < 0, (lo - oi) / d[i], big))
        depth = torch.minimum(depth, t)
    for b in boxes:
        tn = torch.full_like(depth, -1e9)
        tf = big.clone()
        for i in range(3):
            oi = o[:, i].view(V, 1, 1)
            inv = 1.0 / torch.where(d[i] == 0, torch.full_like(d[i], 1e-12), d[i])
            t1, t2 = (float(b[0, i]) - oi) * inv, (float(b[1, i]) - oi) * inv
            tn = torch.maximum(tn, torch.minimum(t1, t2))
            tf = torch.minimum(tf, torch.maximum(t1, t2))
        hit = (tn <= tf) & (tn > 0)
        depth = torch.where(hit, torch.minimum(depth, tn), depth)
    return depth


def make_scene(seed: int, n_frames: int = 32, n_total: int | None = None, H: int = 480, W: int = 640,
               invalid_frac: float = 0.05, device="cpu", frame_indices=None, max_boxes: int = 8) -> Scene:
    """Scene `seed` with `n_frames` frames.  `n_total` is the length of the full capture the
    frames are drawn from (uniform np.linspace picks, llava/video_utils.py:188, unless
    frame_indices is given); n_total == n_frames keeps every frame (MC candidate sets)."""
    rng, room, boxes, axis_align = scene_layout(seed, max_boxes)
    n_total = n_total or n_frames
    c2w_all = camera_loop(rng, room, n_total)
    if frame_indices is None:
        frame_indices = np.linspace(0, n_total - 1, n_frames, dtype=int)
    c2w = c2w_all[np.asarray(frame_indices)]
    scale = W / 640.0  # smaller test images keep the field of view
    fx, fy, cx, cy = FX * scale, FY * scale, (W - 1) / 2.0, (H - 1) / 2.0
    z = render_depth(c2w, room, boxes, H, W, device, fx, fy, cx, cy)
    mm = torch.clamp(torch.round(z * 1000.0), 0, 65535).to(torch.int32)
    if invalid_frac > 0:
        h = _pixel_hash(seed, len(c2w), H, W, device)
        mm = torch.where(h < int(invalid_frac * 65536), torch.zeros_like(mm), mm)
    depth = mm.to(torch.uint16).view(torch.int16) if hasattr(torch, "uint16") else (mm & 0xFFFF).to(torch.int16)
    # sensor pose such that axis_align @ pose == aligned cam2world, as on disk in ScanNet
    sensor = np.linalg.inv(axis_align)[None] @ c2w
    poses32 = torch.tensor((axis_align[None] @ sensor).astype(np.float32), device=device)
    poses32[:, 3, :] = torch.tensor([0.0, 0.0, 0.0, 1.0], device=device)
    Kmat = np.eye(4)
    Kmat[0, 0], Kmat[1, 1], Kmat[0, 2], Kmat[1, 2] = fx, fy, cx, cy
    intr = torch.tensor(np.repeat(Kmat[None], len(c2w), 0).astype(np.float32), device=device)
    return Scene(seed, depth.contiguous(), poses32.contiguous(), intr.contiguous(), room, boxes, axis_align, sensor)


def mesh_voxels(room: np.ndarray, boxes: np.ndarray, seed: int, voxel_size: float = 0.1, step: float = 0.02,
                drop_frac: float = 0.10) -> np.ndarray:
    """Mesh universe pc_voxel: unique round(p / voxel) of points sampled every `step` metres
    (+-5 mm jitter) on every surface, a random `drop_frac` of voxels removed (scanner holes).
    Division is float32 IEEE like VideoProcessor.discrete_point (llava/video_utils.py:348-358).
    Returns (n_pc,3) int32 in a seeded shuffled order (a Python set has no meaningful order)."""
    rng = np.random.default_rng(5000 + seed)
    pts = []

    def faces(lo, hi):
        for ax in range(3):
            u_ax, v_ax = [a for a in range(3) if a != ax]
            us = np.arange(lo[u_ax], hi[u_ax] + 1e-9, step)
            vs = np.arange(lo[v_ax], hi[v_ax] + 1e-9, step)
            uu, vv = np.meshgrid(us, vs, indexing="ij")
            for val in (lo[ax], hi[ax]):
                p = np.empty(uu.shape + (3,))
                p[..., ax], p[..., u_ax], p[..., v_ax] = val, uu, vv
                pts.append(p.reshape(-1, 3))

    faces(room[0], room[1])
    for b in boxes:
        faces(b[0], b[1])
    p = np.concatenate(pts).astype(np.float32)
    p = p + rng.uniform(-0.005, 0.005, p.shape).astype(np.float32)
    ids = np.rint(p / np.float32(voxel_size)).astype(np.int32)
    ids = np.unique(ids, axis=0)
    keep = rng.random(len(ids)) >= drop_frac
    ids = ids[keep]
    return np.ascontiguousarray(ids[rng.permutation(len(ids))])


def make_tokens(seed: int, V: int, n_tok: int, E: int, device="cpu") -> torch.Tensor:
    """N(0,1) bf16 vision tokens for the fused add (seed 2000+seed)."""
    g = torch.Generator(device="cpu").manual_seed(2000 + seed)
    if str(device) == "cpu":
        return torch.randn((V, n_tok, E), generator=g, dtype=torch.float32).to(torch.bfloat16)
    gd = torch.Generator(device=device).manual_seed(2000 + seed)
    return torch.randn((V, n_tok, E), generator=gd, dtype=torch.float32, device=device).to(torch.bfloat16)

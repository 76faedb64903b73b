"""TEST / BASELINE INFRASTRUCTURE ONLY -- torch-CPU port of the reference's front end.

The reference (pure Python) runs this path as a sequence of ATen CPU operators inside DataLoader
workers plus a handful of tiny kernels in the model forward.  /root/reference cannot travel to the
GPU box, so `bench.py`'s cpu_baseline and `--impl reference` legs time this port instead: the same
operator sequence (ATen bmm / elementwise / cv2.resize / avg_pool2d / sin / cos), the same
intermediate tensors, all host threads torch will use.  It is also a second checker: pinned
bit-for-bit against the reference in tests/test_oracle_golden.py::test_torch_port_*.

Nothing under video-3d-llm_b200/ imports this module.
"""
from __future__ import annotations

import cv2
import numpy as np
import torch
import torch.nn.functional as F


def unproject(intrinsics: torch.Tensor, poses: torch.Tensor, depths: torch.Tensor) -> torch.Tensor:
    """llava/video_utils.py:38-68: pixel grid -> camera rays scaled by depth/1000 -> homogeneous
    points -> bmm with the pose -> divide by w.  Materialises the same intermediates as upstream."""
    n, h, w = depths.shape
    rows, cols = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    u = cols.reshape(1, h * w).repeat(n, 1)
    v = rows.reshape(1, h * w).repeat(n, 1)
    k = {name: intrinsics[:, i, j].reshape(n, 1).repeat(1, h * w) for name, (i, j) in
         {"fx": (0, 0), "fy": (1, 1), "cx": (0, 2), "cy": (1, 2)}.items()}
    z = depths.reshape(n, h * w) / 1000
    x = (u - k["cx"]) * z / k["fx"]
    y = (v - k["cy"]) * z / k["fy"]
    cam = torch.stack((x, y, z, torch.ones_like(x)), dim=-1)
    world = torch.bmm(poses, cam.transpose(1, 2)).transpose(1, 2)
    world = world[..., :3] / world[..., 3:4]
    return world.reshape(n, h, w, 3)


def boundary(world: torch.Tensor) -> torch.Tensor:
    """llava/video_utils.py:273-277: six separate reductions with .item()."""
    flat = world.reshape(-1, 3)
    vals = []
    for a in range(3):
        vals += [flat[:, a].min().item(), flat[:, a].max().item()]
    return torch.tensor(vals)


def resize_center_crop(world: torch.Tensor, crop: int = 384) -> torch.Tensor:
    """llava/video_utils.py:296-321: per-frame cv2 nearest resize to (W*crop/H, crop), centre crop,
    np.stack, from_numpy."""
    n, h, w, _ = world.shape
    new_w = int(w * (crop / h))
    left = (new_w - crop) // 2
    frames = [cv2.resize(f.numpy(), (new_w, crop), interpolation=cv2.INTER_NEAREST)[:, left:left + crop, :] for f in world]
    return torch.from_numpy(np.stack(frames))


def average_coordinate_in_patch(coords: torch.Tensor, patch: int = 27) -> torch.Tensor:
    """llava/model/llava_arch.py:213-223."""
    c = coords[:, :-6, :-6, :].permute(0, 3, 1, 2)
    return F.avg_pool2d(c, kernel_size=patch, stride=patch).permute(0, 2, 3, 1)


def discrete_coords(coords: torch.Tensor, lo=(-15, -15, -5), hi=(15, 15, 5), voxel: float = 0.1) -> torch.Tensor:
    """llava/model/llava_arch.py:259-272."""
    lo_t, hi_t = torch.tensor(lo), torch.tensor(hi)
    c = torch.minimum(torch.maximum(coords, lo_t), hi_t)
    return ((c - lo_t) / voxel).round()


def pe_sine3d(x: torch.Tensor, embedding_size: int, temperature: float = 10000) -> torch.Tensor:
    """llava/model/position_encoding.py:17-49 for n_points == 1 and even num_feats."""
    nf = embedding_size // 3
    dim_t = torch.arange(nf, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / nf)
    parts = []
    for a in range(3):
        pos = x[:, :, a, None] / dim_t
        parts.append(torch.stack((pos[:, :, 0::2].sin(), pos[:, :, 1::2].cos()), dim=3).flatten(2))
    pos = torch.cat(parts, dim=2)
    out = torch.zeros((x.shape[0], x.shape[1], embedding_size), dtype=x.dtype)
    out[:, :, :pos.shape[2]] = pos
    return out


def frontend(depths: torch.Tensor, poses: torch.Tensor, intrinsics: torch.Tensor, patch: int, embedding_size: int,
             tokens: torch.Tensor | None = None, with_boundary: bool = False):
    """The chain a1 -> a3(iii) -> a4 -> a5 -> a6 -> a7 of SURVEY.md section 8 as the reference runs it
    (fp32 coordinates, PE cast to the token dtype before the add).  with_boundary adds the unused
    scene-AABB reductions of video_utils.py:273-277 ("as shipped")."""
    world = unproject(intrinsics.float(), poses.float(), depths.float())
    if with_boundary:
        boundary(world)
    coords = resize_center_crop(world)
    pooled = average_coordinate_in_patch(coords, patch)
    disc = discrete_coords(pooled)
    pe = pe_sine3d(disc.flatten(1, 2), embedding_size)
    if tokens is not None:
        return tokens + pe.to(tokens.dtype), pooled, disc
    return pe.to(torch.bfloat16), pooled, disc


# ---- projector tail ---------------------------------------------------------------------------
def get_2dpool(image_feature: torch.Tensor, stride: int = 2, mode: str = "bilinear") -> torch.Tensor:
    """llava/model/llava_arch.py:191-210 with num_patches_per_side = sqrt(tokens)."""
    import math

    num_frames, num_tokens, num_dim = image_feature.shape
    side = int(math.isqrt(num_tokens))
    x = image_feature.view(num_frames, side, side, -1).permute(0, 3, 1, 2).contiguous()
    if mode == "average":
        x = torch.nn.functional.avg_pool2d(x, stride)
    elif mode == "max":
        x = torch.nn.functional.max_pool2d(x, stride)
    elif mode == "bilinear":
        h, w = x.shape[2:]
        x = torch.nn.functional.interpolate(x, size=[math.ceil(h / stride), math.ceil(w / stride)], mode="bilinear")
    else:
        raise ValueError(f"Unexpected mm_spatial_pool_mode: {mode}")
    return x.permute(0, 2, 3, 1).reshape(num_frames, -1, num_dim)


def add_token_per_grid(image_feature: torch.Tensor, image_newline: torch.Tensor) -> torch.Tensor:
    """llava_arch.py:307-328 (add_faster_video off)."""
    import math

    resize_h = int(math.sqrt(image_feature.shape[1]))
    num_frames = image_feature.shape[0]
    x = image_feature.view(num_frames, 1, resize_h, resize_h, -1).permute(4, 0, 2, 1, 3).contiguous()
    x = x.flatten(1, 2).flatten(2, 3)
    x = torch.cat((x, image_newline[:, None, None].expand(*x.shape[:-1], 1)), dim=-1)
    return x.flatten(1, 2).transpose(0, 1)


def projector_tail(image_feature: torch.Tensor, coords, embedding_size: int, image_newline, stride: int = 2, mode: str = "bilinear") -> torch.Tensor:
    """get_2dPool (:469) -> + PE(coords) (:506-517) -> add_token_per_grid (:536).  coords: None or (V, Go*Go, 3) in the
    token dtype (the module returns the PE in the dtype of its input, position_encoding.py:48)."""
    x = get_2dpool(image_feature, stride, mode)
    if coords is not None:
        x = x + pe_sine3d(coords.detach(), embedding_size)
    if image_newline is not None:
        x = add_token_per_grid(x, image_newline)
    return x


# ------------------------------------------------------------------------------------------
# maximum-coverage sampler on the host: the CPU baselines of bench.py
# ------------------------------------------------------------------------------------------
def mc_as_written(depths: torch.Tensor, poses: torch.Tensor, intrinsics: torch.Tensor, pc_voxels, voxel_size: float = 0.1,
                  seed: int = 0, max_pick: int = 32):
    """scripts/3d/preprocessing/max_coverage_sampling.py:34-104 as the reference writes it, with CPU tensors in place of
    `.to('cuda')` (:39) and random.seed(seed) right before the greedy loop: unproject every candidate frame at full
    resolution, divide / round / int, `torch.unique(dim=0)` per frame (:47-50), Python tuples and sets, greedy with
    random.choice tie-breaks.  Returns (order, gains, num_all_voxels, num_select_voxels, seconds per stage)."""
    import random
    import time

    t0 = time.perf_counter()
    world_coords = unproject(intrinsics, poses, depths)
    world_coords = world_coords / voxel_size
    world_coords = world_coords.round().int()
    world_coords_discrete = {}
    for f, coords in enumerate(world_coords):
        coords = torch.unique(coords.view(-1, 3), dim=0)
        world_coords_discrete[f] = [tuple(c) for c in coords.tolist()]
    t1 = time.perf_counter()
    pc_voxel = set(tuple(int(t) for t in c) for c in pc_voxels)
    all_voxel = set()
    for f in world_coords_discrete:
        all_voxel.update(world_coords_discrete[f])
    random.seed(seed)
    picked, gains, used = [], [], set()
    frames = dict(world_coords_discrete)
    for _ in range(len(frames)):
        best, ties = -1, []
        for f in frames:
            cur = set(frames[f]) & pc_voxel
            g = len(cur) - len(used & cur)
            if g > best:
                best, ties = g, [f]
            elif g == best:
                ties.append(f)
        f = random.choice(ties)
        used.update(frames[f])
        picked.append(f)
        gains.append(best)
        frames.pop(f)
        if len(picked) >= max_pick:
            break
    t2 = time.perf_counter()
    return picked, gains, len(all_voxel & pc_voxel), len(used & pc_voxel), {"voxelize_s": t1 - t0, "greedy_s": t2 - t1}


def mc_packed_keys(depths: torch.Tensor, poses: torch.Tensor, intrinsics: torch.Tensor, pc_voxels, voxel_size: float = 0.1,
                   seed: int = 0, max_pick: int = 32):
    """The same selection with the obvious host optimisation (clearly labelled as such in the bench): voxel triples packed
    into int64 keys, `np.unique` per frame, coverage as boolean rows over the mesh universe, greedy with numpy reductions.
    Same picks as mc_as_written (the tie-break protocol is kept)."""
    import random
    import time

    import numpy as np

    t0 = time.perf_counter()
    pc = np.asarray(pc_voxels, np.int64).reshape(-1, 3)
    off = 1 << 20
    key = lambda a: ((a[..., 0] + off) << 42) | ((a[..., 1] + off) << 21) | (a[..., 2] + off)  # noqa: E731
    pc_keys = key(pc)
    order_pc = np.argsort(pc_keys)
    pc_sorted = pc_keys[order_pc]
    F = depths.shape[0]
    cover = np.zeros((F, len(pc_keys)), bool)
    for f in range(F):
        w = unproject(intrinsics[f:f + 1], poses[f:f + 1], depths[f:f + 1])
        ids = (w / voxel_size).round().long().view(-1, 3).numpy()
        k = np.unique(key(ids))
        pos = np.searchsorted(pc_sorted, k)
        pos[pos == len(pc_sorted)] = 0
        hit = pc_sorted[pos] == k
        cover[f, order_pc[pos[hit]]] = True
    t1 = time.perf_counter()
    n_all = int(cover.any(0).sum())
    random.seed(seed)
    used = np.zeros(len(pc_keys), bool)
    alive = list(range(F))
    picked, gains = [], []
    while alive and len(picked) < max_pick:
        g = (cover[alive] & ~used).sum(1)
        best = int(g.max())
        ties = [alive[i] for i in np.nonzero(g == best)[0]]
        f = random.choice(ties)
        used |= cover[f]
        picked.append(f)
        gains.append(best)
        alive.remove(f)
    t2 = time.perf_counter()
    return picked, gains, n_all, int(used.sum()), {"voxelize_s": t1 - t0, "greedy_s": t2 - t1}

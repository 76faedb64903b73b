"""TEST INFRASTRUCTURE ONLY -- numpy front for the CPU oracle (oracle/v3d_oracle.c).

The oracle is the *checker* for the CUDA path: only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / `--impl reference` legs may import this module.  The
product package (video-3d-llm_b200/) never does; it fails loudly without its CUDA
library instead of falling back to anything in here.

Parity status: PINNED against outputs of the reference run in the build container
(oracle/gen_golden.py -> tests/golden/*.npz, checked by tests/test_oracle_golden.py).

Every function cites the reference file:line it restates.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libv3d_oracle.so")

DEPTH_U16, DEPTH_I32, DEPTH_F32 = 0, 1, 2
DIV_IEEE, DIV_RECIPROCAL = 0, 1

_lib = None


def build(force: bool = False) -> str:
    """Compile the C oracle with the committed Makefile (gcc, seconds)."""
    if force or not os.path.isfile(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
        os.path.join(_HERE, "v3d_oracle.c")
    ):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.v3d_oracle_num_threads.restype = ctypes.c_int
    return _lib


def num_threads() -> int:
    return int(lib().v3d_oracle_num_threads())


def set_num_threads(n: int) -> None:
    lib().v3d_oracle_set_num_threads(ctypes.c_int(n))


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _depth_code(depth: np.ndarray) -> int:
    if depth.dtype == np.uint16:
        return DEPTH_U16
    if depth.dtype == np.int32:
        return DEPTH_I32
    if depth.dtype == np.float32:
        return DEPTH_F32
    raise TypeError(f"unsupported depth dtype {depth.dtype}")


# ------------------------------------------------------------------------------------------
# geometry
# ------------------------------------------------------------------------------------------
def unproject(intrinsics, poses, depths, depth_scale: float = 1000.0) -> np.ndarray:
    """llava/video_utils.py:38-68.  (V,4,4),(V,4,4),(V,H,W) -> (V,H,W,3) fp32."""
    depths = np.ascontiguousarray(depths)
    V, H, W = depths.shape
    K = _f32(intrinsics).reshape(V, 16)
    P = _f32(poses).reshape(V, 16)
    out = np.empty((V, H, W, 3), np.float32)
    lib().v3d_oracle_unproject(_p(depths), _depth_code(depths), _p(P), _p(K), V, H, W,
                               ctypes.c_float(depth_scale), _p(out))
    return out


def unproject_uvd_f64(u, v, d, intr, pose) -> np.ndarray:
    """llava/video_utils.py:14-28 (`convert_from_uvd`): scalar float64 twin of unproject,
    an independent high-precision known-answer check for single pixels."""
    intr = np.asarray(intr, np.float64)
    pose = np.asarray(pose, np.float64)
    z = float(d) / 1000
    x = (u - intr[0, 2]) * z / intr[0, 0]
    y = (v - intr[1, 2]) * z / intr[1, 1]
    w = pose @ np.array([x, y, z, 1.0])
    return w[:3] / w[3]


def clamp(world: np.ndarray, lo, hi) -> np.ndarray:
    """llava/video_utils.py:232-234 (the 'norm' token): clamp to the mesh AABB."""
    out = _f32(world).copy()
    lib().v3d_oracle_clamp(_p(out), ctypes.c_size_t(out.size // 3), _p(_f32(lo)), _p(_f32(hi)))
    return out


def boundary(world: np.ndarray) -> np.ndarray:
    """llava/video_utils.py:273-277: [xmin,xmax,ymin,ymax,zmin,zmax] over the full-res coords."""
    w = _f32(world)
    out = np.empty(6, np.float32)
    lib().v3d_oracle_boundary(_p(w), ctypes.c_size_t(w.size // 3), _p(out))
    return out


def resample_maps(H: int, W: int, crop_size: int = 384, strategy: str = "center_crop"):
    """Index maps of the cv2.INTER_NEAREST resize + centre crop in
    llava/video_utils.py:291-308.  OpenCV's nearest kernel reads
    src = min(floor(dst * (1 / (dst_size / src_size))), src_size - 1) in double precision.
    Returns (row_map, col_map) int32, each of length crop_size."""
    if strategy == "resize":
        new_h, new_w, top, left = crop_size, crop_size, 0, 0
    elif strategy == "center_crop":
        new_h = crop_size
        new_w = int(W * (crop_size / H))
        left = (new_w - crop_size) // 2
        top = (new_h - crop_size) // 2
    else:
        raise ValueError(strategy)
    ify = 1.0 / (new_h / H)
    ifx = 1.0 / (new_w / W)
    rows = np.array([min(math.floor(r * ify), H - 1) for r in range(top, top + crop_size)], np.int32)
    cols = np.array([min(math.floor(c * ifx), W - 1) for c in range(left, left + crop_size)], np.int32)
    return rows, cols


def resample(world: np.ndarray, row_map: np.ndarray, col_map: np.ndarray) -> np.ndarray:
    """llava/video_utils.py:296-308 as a gather.  (V,H,W,3) -> (V,len(rows),len(cols),3)."""
    w = _f32(world)
    V, H, W, _ = w.shape
    rm = np.ascontiguousarray(row_map, np.int32)
    cm = np.ascontiguousarray(col_map, np.int32)
    out = np.empty((V, len(rm), len(cm), 3), np.float32)
    lib().v3d_oracle_resample(_p(w), V, H, W, _p(rm), _p(cm), len(rm), len(cm), _p(out))
    return out


# ------------------------------------------------------------------------------------------
# patch pooling, discretisation
# ------------------------------------------------------------------------------------------
def average_coordinate_in_patch(coords: np.ndarray, patch_size: int = 27, trim: int = 6) -> np.ndarray:
    """llava/model/llava_arch.py:213-223.  (V,S,S,3) -> (V,G,G,3)."""
    c = _f32(coords)
    V, S = c.shape[0], c.shape[1]
    G = (S - trim) // patch_size
    out = np.empty((V, G, G, 3), np.float32)
    lib().v3d_oracle_pool_avg(_p(c), V, S, trim, patch_size, _p(out))
    return out


def minmax_coordinate_in_patch(coords: np.ndarray, patch_size: int = 27, trim: int = 6) -> np.ndarray:
    """llava/model/llava_arch.py:225-239.  (V,S,S,3) -> (V,G,G,2,3) (min, max)."""
    c = _f32(coords)
    V, S = c.shape[0], c.shape[1]
    G = (S - trim) // patch_size
    out = np.empty((V, G, G, 2, 3), np.float32)
    lib().v3d_oracle_pool_minmax(_p(c), V, S, trim, patch_size, _p(out))
    return out


def sample_n_points(coords: np.ndarray, n_points: int = 9) -> np.ndarray:
    """llava/model/llava_arch.py:241-257.  (V,384,384,3) -> (V,14,14,n,3) ((V,14,14,3) for n=1)."""
    if n_points not in (1, 5, 9):
        raise NotImplementedError
    c = _f32(coords)
    V, S = c.shape[0], c.shape[1]
    out = np.empty((V, 14, 14, n_points, 3), np.float32)
    lib().v3d_oracle_sample_points(_p(c), V, S, n_points, _p(out))
    return out.reshape(V, 14, 14, 3) if n_points == 1 else out


def discrete_coords(coords: np.ndarray, min_xyz=(-15, -15, -5), max_xyz=(15, 15, 5),
                    voxel_size: float = 0.1, div_mode: int = DIV_IEEE) -> np.ndarray:
    """llava/model/llava_arch.py:259-272 (and llava/video_utils.py:348-358)."""
    c = _f32(coords)
    out = np.empty_like(c)
    lib().v3d_oracle_discretize(_p(c), ctypes.c_size_t(c.size // 3), _p(_f32(min_xyz)), _p(_f32(max_xyz)),
                                ctypes.c_float(voxel_size), div_mode, _p(out))
    return out


# ------------------------------------------------------------------------------------------
# position encoding
# ------------------------------------------------------------------------------------------
def dim_t_table(nf: int, temperature: float = 10000.0) -> np.ndarray:
    """llava/model/position_encoding.py:24-25.  torch's fp32 pow is not correctly rounded, so
    the bit-exact table comes from torch itself (torch is host plumbing on both sides)."""
    import torch

    d = torch.arange(nf, dtype=torch.float32)
    return (temperature ** (2 * (d // 2) / nf)).numpy()


def pe_sine3d(x: np.ndarray, embedding_size: int, temperature: float = 10000.0, n_points: int = 1,
              dim_t: np.ndarray | None = None) -> np.ndarray:
    """llava/model/position_encoding.py:17-49.  (B,N,3) or (B,N,n_points,3) -> (B,N,E) fp32."""
    x = _f32(x)
    nf = embedding_size // (3 * n_points)
    if n_points > 1:
        assert x.ndim == 4 and x.shape[2] == n_points
        B, N = x.shape[0], x.shape[1]
    else:
        B, N = x.shape[0], x.shape[1]
    if dim_t is None:
        dim_t = dim_t_table(nf, temperature)
    dim_t = _f32(dim_t)
    out = np.empty((B, N, embedding_size), np.float32)
    lib().v3d_oracle_pe_sine3d(_p(x), ctypes.c_size_t(B * N), n_points, embedding_size, _p(dim_t), nf, _p(out))
    return out


def f32_to_bf16_bits(a: np.ndarray) -> np.ndarray:
    a = _f32(a)
    out = np.empty(a.shape, np.uint16)
    lib().v3d_oracle_f32_to_bf16(_p(a), ctypes.c_size_t(a.size), _p(out))
    return out


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (np.ascontiguousarray(b, np.uint16).astype(np.uint32) << 16).view(np.float32)


def add_pe_bf16(tokens_bf16_bits: np.ndarray, pe_f32: np.ndarray) -> np.ndarray:
    """llava/model/llava_arch.py:515 in bf16: tokens + bf16(PE), fp32 add, RNE to bf16."""
    t = np.ascontiguousarray(tokens_bf16_bits, np.uint16)
    pe = _f32(pe_f32)
    assert t.size == pe.size
    out = np.empty(t.shape, np.uint16)
    lib().v3d_oracle_add_bf16(_p(t), _p(pe), ctypes.c_size_t(t.size), _p(out))
    return out


def pe_mlp(x: np.ndarray, w1, b1, gamma, beta, w2, b2, eps: float = 1e-5) -> np.ndarray:
    """llava/model/position_encoding.py:52-84 (`PositionEmbeddingMLP`, n_points=1):
    Linear(3,512) -> LayerNorm(512) -> ReLU -> Linear(512,E), float64 accumulate (tolerance
    oracle for the tensor-core kernel)."""
    x = np.asarray(x, np.float64)
    h = x @ np.asarray(w1, np.float64).T + np.asarray(b1, np.float64)
    mu = h.mean(-1, keepdims=True)
    var = h.var(-1, keepdims=True)
    h = (h - mu) / np.sqrt(var + eps) * np.asarray(gamma, np.float64) + np.asarray(beta, np.float64)
    h = np.maximum(h, 0.0)
    return (h @ np.asarray(w2, np.float64).T + np.asarray(b2, np.float64)).astype(np.float32)


def frontend(depths, poses, intrinsics, *, patch_size=27, embedding_size=3584, crop_size=384, trim=6,
             strategy="center_crop", pool="avg", discrete=True, min_xyz=(-15, -15, -5), max_xyz=(15, 15, 5),
             voxel_size=0.1, temperature=10000.0, tokens_bf16_bits=None, clamp_lo=None, clamp_hi=None):
    """The whole front end as the reference chains it: unproject (video_utils.py:230) ->
    [norm clamp :232-234] -> resize/crop (:296-308) -> pool (llava_arch.py:396-406) ->
    discretise (:416) -> sin-3D PE (:515) -> bf16 -> add.  Returns a dict of every stage."""
    V, H, W = depths.shape
    world = unproject(intrinsics, poses, depths)
    if clamp_lo is not None:
        world = clamp(world, clamp_lo, clamp_hi)
    rm, cm = resample_maps(H, W, crop_size, strategy)
    coords = resample(world, rm, cm)
    if pool == "avg":
        pooled, n_points = average_coordinate_in_patch(coords, patch_size, trim), 1
    elif pool == "minmax":
        pooled, n_points = minmax_coordinate_in_patch(coords, patch_size, trim), 2
    elif pool in ("sample9", "sample5", "sample1"):
        n_points = int(pool[-1])
        pooled = sample_n_points(coords, n_points)
    else:
        raise ValueError(pool)
    G = pooled.shape[1]
    pts = pooled.reshape(V, G * G, n_points, 3) if n_points > 1 else pooled.reshape(V, G * G, 3)
    q = discrete_coords(pts, min_xyz, max_xyz, voxel_size) if discrete else pts
    pe = pe_sine3d(q, embedding_size, temperature, n_points)
    out = {"world": world, "coords": coords, "pooled": pooled, "pe_coords": q, "pe_f32": pe,
           "pe_bf16": f32_to_bf16_bits(pe)}
    if tokens_bf16_bits is not None:
        out["tokens_out"] = add_pe_bf16(tokens_bf16_bits, pe)
    return out


# ------------------------------------------------------------------------------------------
# maximum-coverage sampler
# ------------------------------------------------------------------------------------------
def mc_voxel_ids(world: np.ndarray, voxel_size: float = 0.1, div_mode: int = DIV_IEEE) -> np.ndarray:
    """scripts/3d/preprocessing/max_coverage_sampling.py:44-45.  (...,3) fp32 -> (...,3) int32."""
    w = _f32(world)
    out = np.empty(w.shape, np.int32)
    lib().v3d_oracle_mc_voxel_ids(_p(w), ctypes.c_size_t(w.size // 3), ctypes.c_float(voxel_size), div_mode, _p(out))
    return out


def mc_bitsets(ids: np.ndarray, pc_voxels: np.ndarray) -> np.ndarray:
    """max_coverage_sampling.py:47-50,75: per-frame set(frame voxels) & pc_voxel as bitsets over
    the pc list order.  ids (F,...,3) int32, pc (n_pc,3) int32 -> (F, ceil(n_pc/32)) uint32."""
    ids = np.ascontiguousarray(ids, np.int32)
    F = ids.shape[0]
    n_px = ids[0].size // 3 if F else 0
    pc = np.ascontiguousarray(pc_voxels, np.int32).reshape(-1, 3)
    words = (len(pc) + 31) // 32
    out = np.zeros((F, words), np.uint32)
    rc = lib().v3d_oracle_mc_bitsets(_p(ids), F, ctypes.c_size_t(n_px), _p(pc), len(pc), _p(out))
    if rc != 0:
        raise MemoryError
    return out


def mc_greedy(bitsets: np.ndarray, max_pick: int = 32, seed: int = 0):
    """max_coverage_sampling.py:68-104 with random.seed(seed) right before the loop.
    Returns (order[n], gains[n], n_all, n_sel)."""
    bs = np.ascontiguousarray(bitsets, np.uint32)
    F, words = bs.shape
    order = np.empty(max_pick, np.int32)
    gains = np.empty(max_pick, np.int32)
    counts = np.empty(3, np.int32)
    lib().v3d_oracle_mc_greedy(_p(bs), F, words, max_pick, ctypes.c_uint64(seed), _p(order), _p(gains), _p(counts))
    n = int(counts[0])
    return order[:n].copy(), gains[:n].copy(), int(counts[1]), int(counts[2])


def mc_randbelow_seq(seed: int, ns) -> np.ndarray:
    ns = np.ascontiguousarray(ns, np.uint32)
    out = np.empty(len(ns), np.uint32)
    lib().v3d_oracle_mt_randbelow_seq(ctypes.c_uint64(seed), _p(ns), len(ns), _p(out))
    return out


def mc_sample_scene(depths, poses, intrinsics, pc_voxels, voxel_size=0.1, max_pick=32, seed=0,
                    div_mode=DIV_IEEE):
    """Whole MC pipeline for one scene (max_coverage_sampling.py:34-104): unproject every
    candidate frame at full resolution, voxelise, restrict to the mesh universe, greedy."""
    world = unproject(intrinsics, poses, depths)
    ids = mc_voxel_ids(world, voxel_size, div_mode)
    bs = mc_bitsets(ids, pc_voxels)
    return mc_greedy(bs, max_pick, seed) + (bs,)


def mc_python_sets(ids: np.ndarray, pc_voxels, max_pick: int = 32, seed: int = 0):
    """Pure-Python restatement of max_coverage_sampling.py:47-104 with real `set`s and the real
    `random` module -- slow, for small cases; cross-checks the bitset/MT19937 C path."""
    import random

    F = ids.shape[0]
    frames = {}
    for f in range(F):
        uniq = np.unique(ids[f].reshape(-1, 3), axis=0)
        frames[f] = [tuple(int(t) for t in c) for c in uniq]
    all_voxel = set()
    for f in frames:
        all_voxel.update(frames[f])
    pc = set(tuple(int(t) for t in c) for c in np.asarray(pc_voxels).reshape(-1, 3))
    random.seed(seed)
    picked, gains, used = [], [], set()
    for _ in range(F):
        best, ties = -1, []
        for f in frames:
            cur = set(frames[f]) & pc
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
    return (np.array(picked, np.int32), np.array(gains, np.int32), len(all_voxel & pc), len(used & pc))


def mc_prefix_cut(order_files, voxel_nums, num_all_voxels, strategy: str, frames_upbound: int = 32):
    """llava/video_utils.py:131-159 (`sample_frame_files_mc`): first frames_upbound picks, cut at
    the first prefix whose cumulative gain reaches ratio*num_all (0.95 / 0.90 by substring, else
    no cut), then sort by frame number."""
    files = list(order_files[:frames_upbound])
    nums = list(voxel_nums[:frames_upbound])
    ratio = 0.95 if "ratio95" in strategy else (0.9 if "ratio90" in strategy else 1.0)
    if ratio != 1.0:
        out, cc = [], 0
        for f, n in zip(files, nums):
            out.append(f)
            cc += n
            if cc >= num_all_voxels * ratio:
                break
        files = out
    files.sort(key=lambda s: int(str(s).split("/")[-1].split(".")[0]))
    return files


# ------------------------------------------------------------------------------------------
# object proposals
# ------------------------------------------------------------------------------------------
def object_patch_masks(coords: np.ndarray, boxes: np.ndarray, object_feature_type: str = "patch14") -> np.ndarray:
    """llava/model/llava_arch.py:361-376.  coords (V,384,384,3), boxes (K,6) centre+size -> bool (K,V,G,G):
    count of pixels with min <= xyz <= max (all three axes) per patch >= int(p*p*ratio)."""
    if "patch27" in object_feature_type:
        p, g, thr = 27, 14, int(27 * 27 * 0.25)
    elif "patch14" in object_feature_type:
        p, g, thr = 14, 27, int(14 * 14 * 0.5)
    else:
        raise NotImplementedError
    c = _f32(coords)[:, :378, :378, :]
    V = c.shape[0]
    c = c.reshape(V, g, p, g, p, 3).transpose(0, 1, 3, 2, 4, 5).reshape(V, g, g, p * p, 3)
    b = _f32(boxes)
    out = np.zeros((len(b), V, g, g), bool)
    for k in range(len(b)):
        half = b[k, 3:] / np.float32(2)
        lo, hi = b[k, :3] - half, b[k, :3] + half
        inside = np.all((lo <= c) & (c <= hi), axis=-1)
        out[k] = inside.sum(axis=3) >= thr
    return out


def object_features(tokens_bf16_bits: np.ndarray, masks: np.ndarray) -> np.ndarray:
    """llava/model/llava_arch.py:479-496: per object the mean (fp32 accumulate, bf16 result) of the selected token
    rows, zeros when none.  tokens (T,C) bf16 bits, masks (K,T) -> (K,C) bf16 bits."""
    t = bf16_bits_to_f32(tokens_bf16_bits).astype(np.float64)
    m = np.asarray(masks).reshape(len(masks), -1)
    out = np.zeros((len(m), t.shape[1]), np.float32)
    for k in range(len(m)):
        if m[k].any():
            out[k] = t[m[k]].mean(axis=0).astype(np.float32)
    return f32_to_bf16_bits(out)

"""Mirror of the coordinate helpers of LlavaMetaForCausalLM (llava/model/llava_arch.py:213-272)
on the CUDA library: same names, arguments and output shapes/dtypes.

Each helper accepts either a materialised coordinate tensor (V,S,S,3) on a CUDA device, as the
reference passes it, or a `LazyWorldCoords` handle (video_utils.py) over raw depth/pose/intrinsics,
in which case the pooled coordinates come straight out of the fused kernel and the
(V,384,384,3) tensor is never built.
"""
from __future__ import annotations

import torch

from . import cabi

_POOL = {"avg": cabi.POOL_AVG, "minmax": cabi.POOL_MINMAX, "sample9": cabi.POOL_SAMPLE9,
         "sample5": cabi.POOL_SAMPLE5, "sample1": cabi.POOL_SAMPLE1}


def _pool(world_coords, patch_size: int, mode: str, n_points: int, trim: int = 6):
    from .video_utils import LazyWorldCoords

    if isinstance(world_coords, LazyWorldCoords):
        return world_coords.pooled(patch_size, mode, trim)
    cabi.require_cuda(world_coords, "world_coords")
    lib = cabi.load()
    V, H, W, D = world_coords.size()
    if H != W or D != 3:
        raise ValueError(f"world_coords must be (V,S,S,3), got {tuple(world_coords.shape)}")
    grid = (H - trim) // patch_size
    src = world_coords.detach().to(torch.float32).contiguous()
    shape = (V, grid, grid, 3) if n_points == 1 else (V, grid, grid, n_points, 3)
    out = torch.empty(shape, dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        cabi.check(lib.v3d_pool_coords(src.data_ptr(), V, H, patch_size, grid, _POOL[mode], out.data_ptr(),
                                       cabi.stream_ptr(src.device)), "v3d_pool_coords")
    return out.to(world_coords.dtype)


def average_coordinate_in_patch(world_coords, patch_size=27):
    """llava/model/llava_arch.py:213-223: (V,384,384,3) -> (V,G,G,3), G = 378 // patch_size."""
    return _pool(world_coords, patch_size, "avg", 1)


def minmax_coordinate_in_patch(world_coords, patch_size=27):
    """llava/model/llava_arch.py:225-239: -> (V,G,G,2,3) with [...,0,:] the min and [...,1,:] the max."""
    return _pool(world_coords, patch_size, "minmax", 2)


def sample_n_points(world_coords, n_points=9):
    """llava/model/llava_arch.py:241-257: -> (V,14,14,n,3) ((V,14,14,3) for n_points == 1)."""
    if n_points not in (9, 5, 1):
        raise NotImplementedError
    return _pool(world_coords, 27, f"sample{n_points}", n_points)


def discrete_coords(world_coords, xyz_min=None, *, voxel_size=0.1, min_xyz_range=(-15, -15, -5), max_xyz_range=(15, 15, 5),
                    div_mode: int = cabi.DIV_IEEE):
    """llava/model/llava_arch.py:259-272: round((clamp(c, min, max) - min) / voxel_size); the result
    stays floating point and detached; `xyz_min` is accepted and ignored like upstream.  On a bf16 / fp16 tensor every
    step is rounded to that dtype like ATen's element-wise kernels do; a deferred LazyPooled stays deferred."""
    from .lazy import LazyPooled

    if isinstance(world_coords, LazyPooled) and world_coords.disc is None and min_xyz_range is not None and max_xyz_range is not None:
        return world_coords.discretized(voxel_size, min_xyz_range, max_xyz_range, div_mode)
    if isinstance(world_coords, LazyPooled):
        world_coords = world_coords.materialize()
    cabi.require_cuda(world_coords, "world_coords")
    lib = cabi.load()
    if world_coords.shape[-1] != 3:
        raise ValueError("last dimension must be 3")
    src = world_coords.detach().to(torch.float32).contiguous()
    out = torch.empty_like(src)
    has_range = min_xyz_range is not None and max_xyz_range is not None
    lo = cabi.host_floats(min_xyz_range) if has_range else None
    hi = cabi.host_floats(max_xyz_range) if has_range else None
    with torch.cuda.device(src.device):
        cabi.check(lib.v3d_discretize(src.data_ptr(), src.numel() // 3, lo, hi, int(has_range), float(voxel_size), div_mode,
                                      cabi.dtype_code(world_coords.dtype), out.data_ptr(), cabi.stream_ptr(src.device)), "v3d_discretize")
    return out.to(world_coords.dtype)


_OBJECT_GEOMETRY = {"patch14": (14, 27, int(14 * 14 * 0.5)), "patch27": (27, 14, int(27 * 27 * 0.25))}


def object_patch_masks(world_coords, object_boxes, object_feature_type: str = "patch14"):
    """llava/model/llava_arch.py:351-378 for all objects at once.  world_coords (V,384,384,3) (tensor or lazy
    handle), object_boxes (K,6) centre+size -> bool (K,V,G,G): patches with enough pixels inside each box
    ("patch14": 27x27 grid, >= 98 of 196 pixels; "patch27": 14x14 grid, >= 182 of 729 pixels)."""
    from .video_utils import LazyWorldCoords

    for key, geom in _OBJECT_GEOMETRY.items():
        if key in object_feature_type:
            patch, grid, thr = geom
            break
    else:
        raise NotImplementedError(object_feature_type)
    lib = cabi.load()
    if isinstance(world_coords, LazyWorldCoords) and world_coords._full is None and not world_coords.batched:
        # straight from depth: the (V,S,S,3) tensor the reference slices at :360-373 is never built
        from .video_utils import _device_for

        h = world_coords
        dev = _device_for(h.depth)
        depth = h.depth.to(dev).contiguous()
        poses, intr = h.poses.to(dev, torch.float32).contiguous(), h.intrinsics.to(dev, torch.float32).contiguous()
        V, H, W = depth.shape
        rm, cm = h._maps(dev)
        if patch * grid > rm.numel():
            raise ValueError(f"patch*grid={patch * grid} exceeds the {rm.numel()}-pixel crop")
        boxes = object_boxes.detach().to(dev, torch.float32).contiguous()
        K = boxes.shape[0]
        masks = torch.zeros((K, V, grid, grid), dtype=torch.uint8, device=dev)
        lo_hi = None if h.clamp is None else cabi.host_floats(list(h.clamp[0]) + list(h.clamp[1]))
        with torch.cuda.device(dev):
            cabi.check(lib.v3d_object_patch_masks_depth(depth.data_ptr(), cabi.depth_code(depth), poses.data_ptr(), intr.data_ptr(), V, H, W,
                                                        1000.0, rm.data_ptr(), cm.data_ptr(), lo_hi, cabi.dtype_code(h.dtype), patch, grid,
                                                        boxes.data_ptr(), K, thr, masks.data_ptr(), cabi.stream_ptr(dev)),
                       "v3d_object_patch_masks_depth")
        return masks.bool()
    if isinstance(world_coords, LazyWorldCoords):
        world_coords = world_coords.materialize()
    cabi.require_cuda(world_coords, "world_coords")
    V, S = world_coords.shape[0], world_coords.shape[1]
    src = world_coords.detach().to(torch.float32).contiguous()
    boxes = object_boxes.detach().to(src.device, torch.float32).contiguous()
    K = boxes.shape[0]
    masks = torch.zeros((K, V, grid, grid), dtype=torch.uint8, device=src.device)
    with torch.cuda.device(src.device):
        cabi.check(lib.v3d_object_patch_masks(src.data_ptr(), V, S, patch, grid, boxes.data_ptr(), K, thr, masks.data_ptr(),
                                              cabi.stream_ptr(src.device)), "v3d_object_patch_masks")
    return masks.bool()


def object_features(token_features, masks):
    """llava/model/llava_arch.py:479-496: mean token feature of every object's patches, zeros for objects without
    any.  token_features (V,T,C) or (V*T,C) bf16, masks bool/uint8 (K,V,G,G) with G*G == T -> (K,C) bf16."""
    cabi.require_cuda(token_features, "token_features")
    lib = cabi.load()
    C = token_features.shape[-1]
    feats = token_features.detach().reshape(-1, C).to(torch.bfloat16).contiguous()
    K = masks.shape[0]
    m = masks.reshape(K, -1).to(torch.uint8).contiguous()
    if m.shape[1] != feats.shape[0]:
        raise ValueError(f"masks cover {m.shape[1]} tokens, features have {feats.shape[0]}")
    out = torch.empty((K, C), dtype=torch.bfloat16, device=feats.device)
    with torch.cuda.device(feats.device):
        cabi.check(lib.v3d_masked_token_mean(feats.data_ptr(), feats.shape[0], C, m.data_ptr(), K, out.data_ptr(),
                                             cabi.stream_ptr(feats.device)), "v3d_masked_token_mean")
    return out.to(token_features.dtype)


def parse_world_position_embedding_type(kind: str) -> dict:
    """Substring semantics of llava/model/llava_arch.py:51-63, 395-429."""
    if "avg" in kind:
        pool, n_points = "avg", 1
    elif "sample9" in kind:
        pool, n_points = "sample9", 9
    elif "sample5" in kind:
        pool, n_points = "sample5", 5
    elif "sample1" in kind:
        pool, n_points = "sample1", 1
    elif "minmax" in kind:
        pool, n_points = "minmax", 2
    else:
        pool, n_points = None, 1
    return {"pool": pool, "n_points": n_points, "discrete": "discrete" in kind, "sin3d": "sin3d" in kind,
            "mlp": "mlp" in kind, "mrope": "mrope" in kind}

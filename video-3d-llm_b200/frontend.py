"""The fused front end: raw depth + pose + intrinsics -> 3D position encoding (+ vision tokens).

One call = one launch of the fused kernel (v3d_frontend_pe, include/v3d.h): it replaces the chain
  VideoProcessor.calculate_world_coords / preprocess   (llava/video_utils.py:196-238, 291-308)
  average_coordinate_in_patch / discrete_coords        (llava/model/llava_arch.py:213-272, 396-420)
  world_position_embedding(...) and the add            (llava/model/position_encoding.py:17-49, llava_arch.py:515)
without ever materialising world coordinates in HBM.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from dataclasses import dataclass

import torch

from . import cabi, geometry
from .coords import _POOL, parse_world_position_embedding_type
from .position_encoding import get_tables


@dataclass
class FrontEndConfig:
    H: int = 480
    W: int = 640
    crop_size: int = 384
    strategy: str = "center_crop"                 # llava/video_utils.py:293-308
    patch_size: int = 27                          # 27 -> 14x14 grid (reference default); 14 -> 27x27 grid
    trim: int = 6                                 # the [:-6, :-6] of llava_arch.py:217
    world_position_embedding_type: str = "avg-discrete-sin3d"   # scripts/3d/train/train_multi.sh:82
    embedding_size: int = 3584
    temperature: float = 10000
    voxel_size: float = 0.1
    min_xyz_range: tuple = (-15, -15, -5)
    max_xyz_range: tuple = (15, 15, 5)
    depth_scale: float = 1000.0
    dtype: torch.dtype = torch.bfloat16
    div_mode: int = cabi.DIV_IEEE

    @property
    def grid(self) -> int:
        return (self.crop_size - self.trim) // self.patch_size


class FrontEnd:
    """Callable bound to one configuration and device; caches maps, dim_t and the PE table."""

    #: True forces the reference's operator sequence for every patch mean (v3d_frontend_desc.exact_pooling); the results
    #: are identical either way -- this is a switch for tests and profiling
    exact_pooling = False

    def __init__(self, config: FrontEndConfig | None = None, device=None, **overrides):
        cfg = dataclasses.replace(config or FrontEndConfig(), **overrides)
        self.cfg = cfg
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.device.type != "cuda":
            raise RuntimeError("FrontEnd needs a CUDA device: this path has no CPU implementation")
        self.lib = cabi.load()
        self.flags = parse_world_position_embedding_type(cfg.world_position_embedding_type)
        if self.flags["pool"] is None:
            raise ValueError(f"{cfg.world_position_embedding_type!r} names no pooling (avg / sample9 / sample5 / sample1 / minmax)")
        self.n_points = self.flags["n_points"]
        self.row_map, self.col_map = geometry.resample_maps(cfg.H, cfg.W, cfg.crop_size, cfg.strategy, self.device)
        _, cols = geometry.resample_maps_host(cfg.H, cfg.W, cfg.crop_size, cfg.strategy)
        used = cols[: cfg.grid * cfg.patch_size]
        self.src_col_min, self.src_col_max = min(used), max(used)
        self.tables = None
        if self.flags["sin3d"]:
            rows = 1
            if self.flags["discrete"]:
                rows = 1 + max(int(round((hi - lo) / cfg.voxel_size)) for lo, hi in zip(cfg.min_xyz_range, cfg.max_xyz_range))
            self.tables = get_tables(cfg.embedding_size, self.n_points, cfg.temperature, cfg.dtype, self.device, max(rows, 2))
        d = cabi.FrontendDesc()
        d.depth_scale = cfg.depth_scale
        d.row_map, d.col_map = self.row_map.data_ptr(), self.col_map.data_ptr()
        d.src_col_min, d.src_col_max = self.src_col_min, self.src_col_max
        d.patch, d.grid, d.pool_mode = cfg.patch_size, cfg.grid, _POOL[self.flags["pool"]]
        d.discrete = int(self.flags["discrete"])
        d.min_xyz, d.max_xyz = cabi.float3(cfg.min_xyz_range), cabi.float3(cfg.max_xyz_range)
        d.voxel_size, d.div_mode = cfg.voxel_size, cfg.div_mode
        if self.tables is not None:
            d.pe_mode, d.E, d.dtype = cabi.PE_SINE3D, cfg.embedding_size, cabi.dtype_code(cfg.dtype)
            d.dim_t, d.table, d.table_rows = self.tables.dim_t.data_ptr(), self.tables.table.data_ptr(), self.tables.rows
        self._desc0 = bytes(d)

    @property
    def tokens_per_frame(self) -> int:
        return self.cfg.grid * self.cfg.grid

    def upload_window(self, esize: int = 2):
        """(x_off_bytes, width_bytes, row_period, rows_keep) of the depth samples this geometry reads: the 16-byte
        aligned column window and, when the row map has a regular keep-k-of-p structure (ScanNet: 4 of 5), that
        structure; (.., 1, 1) = every row."""
        cfg = self.cfg
        lo = (self.src_col_min * esize) & ~15
        hi = min(((self.src_col_max + 1) * esize + 15) & ~15, cfg.W * esize)
        rows, _ = geometry.resample_maps_host(cfg.H, cfg.W, cfg.crop_size, cfg.strategy)
        used = set(rows[: cfg.grid * cfg.patch_size])
        period, keep = 1, 1
        for per in range(2, 9):
            if cfg.H % per:
                continue
            k = max(r % per for r in used) + 1
            if k < per:
                period, keep = per, k
                break
        return lo, hi - lo, period, keep

    def upload(self, depth_host: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """Stream-ordered host -> device copy of only the depth samples the kernel reads (v3d_upload_depth_window):
        58 % of the bytes for ScanNet-shaped frames.  depth_host: pinned (V,H,W) uint16/int16-bits/int32/float32.
        The rest of the returned device tensor is unspecified and never read by __call__."""
        if depth_host.is_cuda or not depth_host.is_pinned():
            raise ValueError("upload() takes a pinned host tensor")
        depth_host = depth_host.contiguous()
        V, H, W = depth_host.shape
        if (H, W) != (self.cfg.H, self.cfg.W):
            raise ValueError(f"depth is {H}x{W}, front end was configured for {self.cfg.H}x{self.cfg.W}")
        if out is None:
            out = torch.empty(depth_host.shape, dtype=depth_host.dtype, device=self.device)
        es = depth_host.element_size()
        x_off, width, period, keep = self.upload_window(es)
        with torch.cuda.device(self.device):
            cabi.check(self.lib.v3d_upload_depth_window(depth_host.data_ptr(), out.data_ptr(), W * es, V * H, x_off, width, period, keep,
                                                        cabi.stream_ptr(self.device)), "v3d_upload_depth_window")
        return out

    def packed_shape(self, n_frames: int, esize: int = 2):
        """(n_frames, rows kept per frame, window width in elements) of the packed window layout."""
        x_off, width, period, keep = self.upload_window(esize)
        return n_frames, self.cfg.H // period * keep, width // esize

    def pack_window(self, depth_host: torch.Tensor, out: torch.Tensor | None = None, pin: bool = True) -> torch.Tensor:
        """Host-side: keep only the depth samples the kernel reads, contiguously (what a loader that already skips the
        rest produces, e.g. ScenePack.load(window=...)).  (V,H,W) -> (V, rows kept, window width)."""
        V, H, W = depth_host.shape
        es = depth_host.element_size()
        x_off, width, period, keep = self.upload_window(es)
        if out is None:
            out = torch.empty(self.packed_shape(V, es), dtype=depth_host.dtype, pin_memory=pin)
        src = depth_host.view(V, H // period, period, W)[:, :, :keep, x_off // es:(x_off + width) // es]
        out.view(V, H // period, keep, width // es).copy_(src)
        return out

    def upload_packed(self, packed_host: torch.Tensor, staging: torch.Tensor | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
        """Stream-ordered: ONE contiguous host -> device copy of a packed window (pinned), then a device-side strided
        expansion into the (V,H,W) layout __call__ reads.  Samples outside the window are unspecified and never read."""
        if packed_host.is_cuda or not packed_host.is_pinned():
            raise ValueError("upload_packed() takes a pinned host tensor")
        V = packed_host.shape[0]
        es = packed_host.element_size()
        if tuple(packed_host.shape) != self.packed_shape(V, es) or not packed_host.is_contiguous():
            raise ValueError(f"packed depth must be contiguous {self.packed_shape(V, es)}, got {tuple(packed_host.shape)}")
        if staging is None:
            staging = torch.empty(packed_host.shape, dtype=packed_host.dtype, device=self.device)
        if out is None:
            out = torch.empty((V, self.cfg.H, self.cfg.W), dtype=packed_host.dtype, device=self.device)
        with torch.cuda.device(self.device):
            staging.copy_(packed_host, non_blocking=True)
        return self.expand_packed(staging, out)

    def expand_packed(self, staging: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """Device-side strided expansion of a packed window (already on the device) into the (V,H,W) layout, on the
        current stream -- separate from the copy so that a pipeline can keep its copy stream busy with DMAs only."""
        V = staging.shape[0]
        es = staging.element_size()
        x_off, width, period, keep = self.upload_window(es)
        with torch.cuda.device(self.device):
            cabi.check(self.lib.v3d_expand_depth_window(staging.data_ptr(), out.data_ptr(), self.cfg.W * es, V * self.cfg.H, x_off, width,
                                                        period, keep, cabi.stream_ptr(self.device)), "v3d_expand_depth_window")
        return out

    # ---- resampled form: the reference's resize + crop applied to the raw depth by the loader (crop x crop samples per
    # frame, 384 x 384 for the ScanNet recipe: rows of 768 bytes keep the TMA path; the kernel reads the first grid*patch)
    def resampled_shape(self, n_frames: int):
        return n_frames, self.cfg.crop_size, self.cfg.crop_size

    def pack_resampled(self, depth_host: torch.Tensor, out: torch.Tensor | None = None, pin: bool = True) -> torch.Tensor:
        """Host-side nearest-neighbour resize + crop of the raw depth (llava/video_utils.py:291-308 applied to depth instead
        of coordinates): (V,H,W) -> (V,crop,crop).  What a loader that resamples on the host emits (ScenePack.load(
        resample=...), or reads straight from a pack written with a resampled stack); 48 % of the bytes of a ScanNet frame."""
        V = depth_host.shape[0]
        rows, cols = geometry.resample_maps_host(self.cfg.H, self.cfg.W, self.cfg.crop_size, self.cfg.strategy)
        rows_t, cols_t = torch.as_tensor(rows, dtype=torch.long), torch.as_tensor(cols, dtype=torch.long)
        if out is None:
            out = torch.empty(self.resampled_shape(V), dtype=depth_host.dtype, pin_memory=pin)
        torch.index_select(depth_host.index_select(1, rows_t), 2, cols_t, out=out)
        return out

    def scatter_resampled(self, staging: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """Device side: put resampled depth (V,R,C) (already on the device) back at the pixels it was taken from in the
        (V,H,W) layout, on the current stream -- only for callers that need the frame layout (the fused kernel reads the
        resampled form in place, see from_resampled)."""
        V, R, Cc = staging.shape
        if R > self.cfg.crop_size or Cc > self.cfg.crop_size or tuple(out.shape) != (V, self.cfg.H, self.cfg.W) or out.dtype != staging.dtype:
            raise ValueError(f"resampled depth must be at most {self.resampled_shape(V)} and out (V,{self.cfg.H},{self.cfg.W}) of the same dtype")
        staging = staging.contiguous()
        with torch.cuda.device(self.device):
            cabi.check(self.lib.v3d_scatter_resampled_depth(staging.data_ptr(), out.data_ptr(), cabi.depth_code(staging), V, self.cfg.H,
                                                            self.cfg.W, self.row_map.data_ptr(), self.col_map.data_ptr(), R, Cc,
                                                            cabi.stream_ptr(self.device)), "v3d_scatter_resampled_depth")
        return out

    def from_resampled(self, resampled: torch.Tensor, poses, intrinsics, **kw):
        """__call__ on depth the loader already resampled ((V,R,C) with R, C >= grid*patch; pinned host or device): uploaded
        if needed and read in place by the fused kernel (v3d_frontend_desc.depth_resampled) -- no scatter, no scratch."""
        if not resampled.is_cuda:
            resampled = resampled.to(self.device, non_blocking=True)
        return self(resampled, poses.to(self.device, non_blocking=True), intrinsics.to(self.device, non_blocking=True), resampled=True, **kw)

    def upload_bytes(self, n_frames: int, esize: int = 2) -> int:
        x_off, width, period, keep = self.upload_window(esize)
        return n_frames * (self.cfg.H // period) * keep * width

    def __call__(self, depth, poses, intrinsics, tokens=None, out=None, clamp=None, want_coords=False, resampled=False,
                 coord_dtype: torch.dtype = torch.float32, pe: bool = True):
        """depth (V,H,W) uint16/int16-bits/int32/float32, poses & intrinsics (V,4,4) float32, all on the
        device.  tokens: optional (V, G*G, E) in cfg.dtype to add the PE to (may be `out`).
        resampled: depth is (V,R,C) as the loader resampled it (depth[v,r,c] = raw[v,row_map[r],col_map[c]]).
        coord_dtype: the dtype the reference's world_coords tensor would have (fp32 = the oracle's contract; bf16 / fp16
        reproduce the roundings of the Trainer-cast / .half() tensors of llava_arch.py:396-420).
        pe=False: stop after the coordinates (pooling + discretisation only), whatever the configured PE.
        Returns the PE (or tokens + PE) as (V, G*G, E); with want_coords also (pooled, pe_coords)."""
        cfg = self.cfg
        if not (depth.is_cuda and poses.is_cuda and intrinsics.is_cuda):
            for name, t in (("depth", depth), ("poses", poses), ("intrinsics", intrinsics)):
                cabi.require_cuda(t, name)
        V, H, W = depth.shape
        if resampled:
            n = cfg.grid * cfg.patch_size
            if H < n or W < n:
                raise ValueError(f"resampled depth is {H}x{W}, the {cfg.grid}x{cfg.grid} grid of {cfg.patch_size}-pixel patches needs {n}x{n}")
        elif (H, W) != (cfg.H, cfg.W):
            raise ValueError(f"depth is {H}x{W}, front end was configured for {cfg.H}x{cfg.W}")
        if not depth.is_contiguous():
            depth = depth.contiguous()
        if poses.dtype != torch.float32 or not poses.is_contiguous():
            poses = poses.to(torch.float32).contiguous()
        if intrinsics.dtype != torch.float32 or not intrinsics.is_contiguous():
            intrinsics = intrinsics.to(torch.float32).contiguous()
        G, E = cfg.grid, cfg.embedding_size
        # the per-configuration fields are filled once (self._desc0); a call copies the struct and sets what changes
        d = cabi.FrontendDesc.from_buffer_copy(self._desc0)
        d.depth, d.depth_dtype, d.poses, d.intr = depth.data_ptr(), cabi.depth_code(depth), poses.data_ptr(), intrinsics.data_ptr()
        d.n_frames, d.H, d.W = V, H, W
        if clamp is not None:
            d.has_clamp, d.clamp_lo, d.clamp_hi = 1, cabi.float3(clamp[0]), cabi.float3(clamp[1])
        d.depth_resampled, d.coord_dtype = int(bool(resampled)), cabi.dtype_code(coord_dtype)
        d.exact_pooling = int(bool(self.exact_pooling))
        pooled = pe_coords = None
        use_pe = pe and self.flags["sin3d"]
        if want_coords or not use_pe:
            shape = (V, G, G, 3) if self.n_points == 1 else (V, G, G, self.n_points, 3)
            pooled = torch.empty(shape, dtype=torch.float32, device=self.device)
            pe_coords = torch.empty(shape, dtype=torch.float32, device=self.device)
            d.pooled_out, d.pe_coords_out = pooled.data_ptr(), pe_coords.data_ptr()
        if use_pe:
            if out is None:
                out = torch.empty((V, G * G, E), dtype=cfg.dtype, device=self.device)
            elif out.shape != (V, G * G, E) or out.dtype != cfg.dtype or not out.is_contiguous():
                raise ValueError("out must be a contiguous (V, G*G, E) tensor of the configured dtype")
            if tokens is not None:
                if tokens.shape != out.shape or tokens.dtype != cfg.dtype:
                    raise ValueError(f"tokens must be {tuple(out.shape)} {cfg.dtype}, got {tuple(tokens.shape)} {tokens.dtype}")
                if not tokens.is_contiguous():
                    tokens = tokens.contiguous()
                d.tokens_in = tokens.data_ptr()
            d.pe_out = out.data_ptr()
        else:
            d.pe_mode = cabi.PE_NONE
        with cabi.on_device(self.device):
            cabi.check(self.lib.v3d_frontend_pe(C.byref(d), cabi.stream_ptr(self.device)), "v3d_frontend_pe")
        if want_coords:
            return (out if use_pe else None), pooled, pe_coords
        return out if use_pe else pe_coords

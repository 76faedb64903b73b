"""ctypes binding of the C ABI declared in include/v3d.h (libv3d.so, built by build.py).

There is no CPU fallback anywhere in the product path: if the library is missing or a call
fails, these helpers raise.  torch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("V3D_LIB_PATH") or os.path.join(_HERE, "lib", "libv3d.so")  # override: experimental builds

# enums of include/v3d.h
V3D_OK, V3D_E_ARG, V3D_E_ALIGN, V3D_E_UNSUPPORTED, V3D_E_CUDA = 0, -1, -2, -3, -4
DEPTH_U16, DEPTH_I32, DEPTH_F32 = 0, 1, 2
BF16, F16, F32 = 0, 1, 2
DIV_IEEE, DIV_RECIPROCAL = 0, 1
POOL_AVG, POOL_MINMAX, POOL_SAMPLE9, POOL_SAMPLE5, POOL_SAMPLE1 = 0, 1, 2, 3, 4
PE_NONE, PE_SINE3D = 0, 1
TOKEN_POOL_BILINEAR, TOKEN_POOL_AVERAGE, TOKEN_POOL_MAX, TOKEN_POOL_NONE = 0, 1, 2, 3
MC_MAX_PICK = 32
MC_RECORD_INTS = 2 * MC_MAX_PICK + 4

# every symbol include/v3d.h declares (tests check that the library exports them all)
SYMBOLS = (
    "v3d_version", "v3d_last_error", "v3d_selftest_division", "v3d_pool_points", "v3d_unproject", "v3d_upload_depth_window", "v3d_expand_depth_window", "v3d_scatter_resampled_depth", "v3d_host_gather", "v3d_h2d_gather",
    "v3d_boundary", "v3d_pool_coords",
    "v3d_discretize", "v3d_pe_table_bytes", "v3d_pe_table_build", "v3d_pe_sine3d", "v3d_frontend_pe",
    "v3d_mc_hash_slots", "v3d_mc_sample", "v3d_mc_voxelize", "v3d_mc_greedy", "v3d_mc_voxel_ids", "v3d_pe_mlp", "v3d_object_patch_masks", "v3d_object_patch_masks_depth", "v3d_object_patch_masks_depth_bounds", "v3d_masked_token_mean",
    "v3d_token_pool_out_grid", "v3d_pool_pe_grid",
)


class FrontendDesc(C.Structure):
    """struct v3d_frontend_desc"""

    _fields_ = [
        ("depth", C.c_void_p), ("depth_dtype", C.c_int), ("poses", C.c_void_p), ("intr", C.c_void_p),
        ("n_frames", C.c_int), ("H", C.c_int), ("W", C.c_int), ("depth_scale", C.c_float),
        ("row_map", C.c_void_p), ("col_map", C.c_void_p), ("src_col_min", C.c_int), ("src_col_max", C.c_int),
        ("patch", C.c_int), ("grid", C.c_int), ("pool_mode", C.c_int), ("has_clamp", C.c_int),
        ("clamp_lo", C.c_float * 3), ("clamp_hi", C.c_float * 3),
        ("discrete", C.c_int), ("min_xyz", C.c_float * 3), ("max_xyz", C.c_float * 3), ("voxel_size", C.c_float),
        ("div_mode", C.c_int),
        ("pe_mode", C.c_int), ("E", C.c_int), ("dim_t", C.c_void_p), ("table", C.c_void_p), ("table_rows", C.c_int),
        ("dtype", C.c_int), ("tokens_in", C.c_void_p), ("pe_out", C.c_void_p),
        ("pooled_out", C.c_void_p), ("pe_coords_out", C.c_void_p),
        ("depth_resampled", C.c_int), ("coord_dtype", C.c_int), ("exact_pooling", C.c_int),
    ]

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.coord_dtype = F32  # V3D_BF16 is 0: the zero-initialised struct would otherwise ask for bf16 rounding


class PoolDesc(C.Structure):
    """struct v3d_pool_desc"""

    _fields_ = [
        ("tokens", C.c_void_p), ("n_frames", C.c_int), ("src_grid", C.c_int), ("E", C.c_int), ("dtype", C.c_int),
        ("pool_mode", C.c_int), ("stride", C.c_int), ("coords", C.c_void_p), ("n_points", C.c_int), ("dim_t", C.c_void_p),
        ("table", C.c_void_p), ("table_rows", C.c_int), ("newline", C.c_void_p), ("out", C.c_void_p),
    ]


class McScene(C.Structure):
    """struct v3d_mc_scene"""

    _fields_ = [
        ("depth", C.c_void_p), ("poses", C.c_void_p), ("intr", C.c_void_p), ("pc_voxels", C.c_void_p),
        ("n_frames", C.c_int), ("n_pc", C.c_int), ("seed", C.c_uint64),
        ("hash_table", C.c_void_p), ("hash_slots", C.c_int), ("pad_", C.c_int),
        ("bitsets", C.c_void_p), ("record", C.c_void_p),
    ]


class V3DError(RuntimeError):
    pass


_lib = None


def load():
    """Load libv3d.so or raise -- the product never falls back to a CPU implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python video-3d-llm_b200/build.py` "
            "(or __graft_entry__.build()); there is no CPU fallback for this path"
        )
    lib = C.CDLL(LIB_PATH)
    vp, i, f, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    lib.v3d_version.restype = i
    lib.v3d_last_error.restype = C.c_char_p
    lib.v3d_pool_points.argtypes, lib.v3d_pool_points.restype = [i], i
    lib.v3d_selftest_division.argtypes = [f, vp, vp]
    lib.v3d_unproject.argtypes = [vp, i, vp, vp, i, i, i, f, vp, vp, i, i, vp, vp, vp]
    lib.v3d_boundary.argtypes = [vp, i, vp, vp, i, i, i, f, vp, vp, vp]
    lib.v3d_upload_depth_window.argtypes = [vp, vp, sz, sz, sz, sz, sz, sz, vp]
    lib.v3d_expand_depth_window.argtypes = [vp, vp, sz, sz, sz, sz, sz, sz, vp]
    lib.v3d_scatter_resampled_depth.argtypes = [vp, vp, i, i, i, i, vp, vp, i, i, vp]
    lib.v3d_host_gather.argtypes = [vp, vp, sz, i, i]
    lib.v3d_h2d_gather.argtypes = [vp, vp, sz, i, vp]
    lib.v3d_pool_coords.argtypes = [vp, i, i, i, i, i, vp, vp]
    lib.v3d_discretize.argtypes = [vp, sz, vp, vp, i, f, i, i, vp, vp]
    lib.v3d_pe_table_bytes.argtypes, lib.v3d_pe_table_bytes.restype = [i, i, i, i], sz
    lib.v3d_pe_table_build.argtypes = [vp, i, i, i, i, vp, vp]
    lib.v3d_pe_sine3d.argtypes = [vp, sz, i, i, vp, vp, i, i, vp, vp, vp]
    lib.v3d_frontend_pe.argtypes = [C.POINTER(FrontendDesc), vp]
    lib.v3d_mc_hash_slots.argtypes, lib.v3d_mc_hash_slots.restype = [i], i
    lib.v3d_mc_sample.argtypes = [C.POINTER(McScene), i, i, i, i, f, f, i, i, vp, vp]
    lib.v3d_mc_voxelize.argtypes = [C.POINTER(McScene), i, i, i, i, f, f, i, vp, vp]
    lib.v3d_mc_greedy.argtypes = [C.POINTER(McScene), i, i, vp, vp]
    lib.v3d_mc_voxel_ids.argtypes = [vp, sz, f, i, vp, vp]
    lib.v3d_pe_mlp.argtypes = [vp, i, i, vp, vp, vp, vp, f, vp, vp, vp, vp, vp]
    lib.v3d_object_patch_masks.argtypes = [vp, i, i, i, i, vp, i, i, vp, vp]
    lib.v3d_object_patch_masks_depth.argtypes = [vp, i, vp, vp, i, i, i, f, vp, vp, vp, i, i, i, vp, i, i, vp, vp]
    lib.v3d_object_patch_masks_depth_bounds.argtypes = [vp, i, vp, vp, i, i, i, f, vp, vp, vp, i, i, i, vp, i, i, vp, vp]
    lib.v3d_masked_token_mean.argtypes = [vp, i, i, vp, i, vp, vp]
    lib.v3d_token_pool_out_grid.argtypes = [i, i, i]
    lib.v3d_pool_pe_grid.argtypes = [C.POINTER(PoolDesc), vp]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("v3d_version", "v3d_last_error", "v3d_selftest_division", "v3d_pool_points", "v3d_pe_table_bytes", "v3d_mc_hash_slots"):
            fn.restype = i
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc == V3D_OK:
        return
    msg = load().v3d_last_error().decode(errors="replace")
    text = f"{what}: {msg} (code {rc})" if what else f"{msg} (code {rc})"
    if rc in (V3D_E_ARG, V3D_E_ALIGN):
        raise ValueError(text)
    if rc == V3D_E_UNSUPPORTED:
        raise NotImplementedError(text)
    raise V3DError(text)


def stream_ptr(device=None) -> int:
    """Raw handle of torch's current stream on `device` (the stream every launch of this package is ordered on)."""
    if device is None:
        idx = torch.cuda.current_device()
    else:
        idx = device if isinstance(device, int) else torch.device(device).index
        if idx is None:
            idx = torch.cuda.current_device()
    raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)  # a tenth of the cost of building a Stream object per launch
    return int(raw(idx)) if raw is not None else torch.cuda.current_stream(idx).cuda_stream


class on_device:
    """`with torch.cuda.device(dev)` that costs nothing when `dev` already is the current device (the usual case: one
    process per GPU) -- the context manager is several microseconds per launch, as much as a one-scene kernel's launch."""

    __slots__ = ("idx", "ctx")

    def __init__(self, device):
        self.idx = torch.device(device).index
        self.ctx = None

    def __enter__(self):
        if self.idx is not None and self.idx != torch.cuda.current_device():
            self.ctx = torch.cuda.device(self.idx)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            return self.ctx.__exit__(*exc)
        return False


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} must live on a CUDA device: the front end has no CPU path (got device {t.device}); "
            "move the tensor with .cuda() or run the reference implementation instead"
        )


def depth_code(t: torch.Tensor) -> int:
    """Depth tensors: uint16 (or its int16 bit view -- torch.uint16 support is partial), int32, float32."""
    if t.dtype in (torch.int16, getattr(torch, "uint16", torch.int16)):
        return DEPTH_U16
    if t.dtype == torch.int32:
        return DEPTH_I32
    if t.dtype == torch.float32:
        return DEPTH_F32
    raise TypeError(f"unsupported depth dtype {t.dtype}; use uint16/int16 (raw PNG bits), int32 or float32")


def div_mode_default() -> int:
    """Division semantics of the drop-in entry points that replace code the reference runs ON CUDA (the overlay's
    discrete_coords, llava_arch.py:269; the MC script's voxel ids, max_coverage_sampling.py:44): ATen's CUDA kernel
    multiplies by the reciprocal of a scalar divisor, so that is the default; V3D_DIV_MODE=ieee selects torch-CPU
    semantics (the parity target of the CPU oracle and of the golden vectors)."""
    v = os.environ.get("V3D_DIV_MODE", "reciprocal").lower()
    if v not in ("ieee", "reciprocal"):
        raise ValueError(f"V3D_DIV_MODE must be 'ieee' or 'reciprocal', got {v!r}")
    return DIV_IEEE if v == "ieee" else DIV_RECIPROCAL


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.bfloat16:
        return BF16
    if dt == torch.float16:
        return F16
    if dt == torch.float32:
        return F32
    raise TypeError(f"unsupported PE dtype {dt}")


def float3(vals):
    return (C.c_float * 3)(*[float(v) for v in vals])


def host_floats(vals):
    """A ctypes float array (HOST pointer argument) from a python sequence, or None."""
    if vals is None:
        return None
    vals = [float(v) for v in vals]
    return (C.c_float * len(vals))(*vals)

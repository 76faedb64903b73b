"""Deferred values that let the reference's own call structure end in ONE fused kernel launch.

`prepare_inputs_labels_for_multimodal` (llava/model/llava_arch.py:384-429, 506-517) computes, per scene,

    coords = self.average_coordinate_in_patch(world_coords)          :396   (or sample_n_points / minmax, :398-406)
    coords = self.discrete_coords(coords, xyz_min[i])                :416
    image_feat = image_feat + self.get_model().world_position_embedding(coords.flatten(1, 2).detach())    :510-515

With the overlay launcher (launch.py) `world_coords` is a `LazyWorldCoords` handle over raw depth/pose/intrinsics, the
helpers return `LazyPooled` (nothing is launched), the PE module returns `LazyPE`, and the `+` of :515 -- the first
point where the vision tokens meet the coordinates -- launches `v3d_frontend_pe` with `tokens_in` set: unproject,
resample, pool, discretise, PE and add in one pass over the depth and the tokens.  Everything else the method (or
any other caller) does to these objects materialises the tensor the reference would hold at that point, with the
same shape, dtype and values, through the same library.
"""
from __future__ import annotations

import torch

from . import cabi

_F = (torch.float32, torch.bfloat16, torch.float16)


def _unwrap(x):
    if isinstance(x, DeferredTensor):
        return x.materialize()
    if isinstance(x, (list, tuple)):
        return type(x)(_unwrap(v) for v in x)
    if isinstance(x, dict):
        return {k: _unwrap(v) for k, v in x.items()}
    return x


class DeferredTensor:
    """Base of the deferred values: behaves like the tensor `materialize()` returns."""

    def materialize(self) -> torch.Tensor:  # pragma: no cover - abstract
        raise NotImplementedError

    # attribute access, indexing, iteration and torch functions fall back to the real tensor
    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.materialize(), name)

    def __getitem__(self, idx):
        return self.materialize()[idx]

    def __iter__(self):
        return iter(self.materialize())

    def __len__(self):
        return self.shape[0]

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        return func(*_unwrap(args), **_unwrap(kwargs or {}))

    def _binary(name):  # noqa: N805 -- dunders are looked up on the type, so the common ones are spelled out
        def op(self, other):
            return getattr(self.materialize(), name)(_unwrap(other))

        op.__name__ = name
        return op

    for _n in ("__add__", "__radd__", "__sub__", "__rsub__", "__mul__", "__rmul__", "__truediv__", "__rtruediv__", "__eq__", "__ne__",
               "__lt__", "__le__", "__gt__", "__ge__", "__matmul__"):
        locals()[_n] = _binary(_n)
    del _n, _binary
    __hash__ = object.__hash__

    def __neg__(self):
        return -self.materialize()

    # cheap metadata shared by the subclasses (they define .shape / .dtype / .device)
    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def dim(self):
        return len(self.shape)

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def is_cuda(self):
        return self.device.type == "cuda"

    @property
    def requires_grad(self):
        return False

    def detach(self):
        return self


class LazyPooled(DeferredTensor):
    """Result of average_coordinate_in_patch / minmax_coordinate_in_patch / sample_n_points (llava_arch.py:213-257) on a
    `LazyWorldCoords` handle, optionally followed by discrete_coords (:259-272) and `.flatten(1, 2)` (:510-512)."""

    def __init__(self, handle, patch_size: int, mode: str, trim: int, n_points: int, disc=None, flat: bool = False):
        self.handle, self.patch_size, self.mode, self.trim, self.n_points = handle, patch_size, mode, trim, n_points
        self.disc = disc  # None or dict(voxel_size, min_xyz_range, max_xyz_range, div_mode)
        self.flat = flat
        self._val = None

    def _replace(self, **kw):
        a = dict(handle=self.handle, patch_size=self.patch_size, mode=self.mode, trim=self.trim, n_points=self.n_points, disc=self.disc,
                 flat=self.flat)
        a.update(kw)
        return LazyPooled(**a)

    @property
    def grid(self) -> int:
        return (self.handle.crop_size - self.trim) // self.patch_size

    @property
    def shape(self):
        V, G = self.handle.shape[0], self.grid
        lead = (V, G * G) if self.flat else (V, G, G)
        return torch.Size(lead + ((3,) if self.n_points == 1 else (self.n_points, 3)))

    @property
    def dtype(self):
        return self.handle.dtype

    @property
    def device(self):
        return self.handle.device

    def flatten(self, start_dim=0, end_dim=-1):
        if (start_dim, end_dim) == (1, 2) and not self.flat:
            return self._replace(flat=True)
        return self.materialize().flatten(start_dim, end_dim)

    def discretized(self, voxel_size, min_xyz_range, max_xyz_range, div_mode):
        if self.disc is not None:
            raise ValueError("coordinates are already discretised")
        return self._replace(disc=dict(voxel_size=float(voxel_size), min_xyz_range=tuple(float(v) for v in min_xyz_range),
                                       max_xyz_range=tuple(float(v) for v in max_xyz_range), div_mode=int(div_mode)))

    # ---- evaluation
    def frontend(self, device, embedding_size=None, temperature=10000.0, out_dtype=torch.bfloat16):
        """The FrontEnd (cached per configuration and device) that evaluates this value (+ the PE when embedding_size is given)."""
        from . import position_encoding as _pe
        from .frontend import FrontEnd

        h = self.handle
        kind = self.mode + ("-discrete" if self.disc is not None else "") + ("-sin3d" if embedding_size else "")
        disc = self.disc or dict(voxel_size=0.1, min_xyz_range=(-15.0, -15.0, -5.0), max_xyz_range=(15.0, 15.0, 5.0), div_mode=cabi.DIV_IEEE)
        H, W = h.src_hw
        key = (H, W, h.crop_size, h.strategy, self.patch_size, self.trim, kind, embedding_size, temperature, out_dtype,
               disc["voxel_size"], disc["min_xyz_range"], disc["max_xyz_range"], disc["div_mode"], device, _pe._dim_t_source)
        fe = _frontends.get(key)
        if fe is None:
            fe = FrontEnd(device=device, H=H, W=W, crop_size=h.crop_size, strategy=h.strategy, patch_size=self.patch_size, trim=self.trim,
                          world_position_embedding_type=kind, embedding_size=embedding_size or 3584, temperature=temperature,
                          voxel_size=disc["voxel_size"], min_xyz_range=disc["min_xyz_range"], max_xyz_range=disc["max_xyz_range"],
                          dtype=out_dtype, div_mode=disc["div_mode"])
            _frontends[key] = fe
        return fe

    def inputs(self, device):
        """(depth, poses, intrinsics, clamp list, resampled) on `device`; host tensors are copied asynchronously (they are
        pinned when they come out of a DataLoader with pin_memory=True, see LazyWorldCoords.pin_memory)."""
        h = self.handle
        mv = lambda t: t if t.device == device else t.to(device, non_blocking=True)  # noqa: E731
        clamp = [] if h.clamp is None else [float(v) for v in h.clamp[0]] + [float(v) for v in h.clamp[1]]
        return mv(h.depth), mv(h.poses), mv(h.intrinsics), clamp, bool(h.resampled)

    def materialize(self) -> torch.Tensor:
        if self._val is None:
            from . import ops
            from .video_utils import _device_for

            dev = _device_for(self.handle.depth)
            fe = self.frontend(dev)
            depth, poses, intr, clamp, resampled = self.inputs(dev)
            if _dispatch():
                pooled, pe_coords = torch.ops.v3d.frontend_coords(depth, poses, intr, ops.frontend_key(fe), clamp, self.dtype, resampled)
            else:
                _, pooled, pe_coords = fe(depth, poses, intr, clamp=ops._clamp_arg(clamp), want_coords=True, resampled=resampled,
                                          coord_dtype=self.dtype, pe=False)
            val = (pe_coords if self.disc is not None else pooled).to(self.dtype)
            self._val = val.flatten(1, 2) if self.flat else val
        return self._val


_frontends: dict = {}
_ADD_FUNCS = None


def _dispatch(*tensors) -> bool:
    """Route a launch through torch.ops.v3d.* (autograd formula, traceable node) or call the library directly?  The
    dispatcher costs tens of microseconds per call, which is as much as the kernel itself for one 32-frame scene; it is
    only needed when a gradient has to flow or the call is being traced."""
    if torch.compiler.is_compiling():
        return True
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def _add_funcs():
    global _ADD_FUNCS
    if _ADD_FUNCS is None:
        _ADD_FUNCS = {torch.add, torch.Tensor.add, torch.Tensor.__add__, torch.Tensor.__radd__}
    return _ADD_FUNCS


class LazyPE(DeferredTensor):
    """Result of PositionEmbeddingSine3D.forward (llava/model/position_encoding.py:17-49): `tokens + self` is one fused
    launch (llava_arch.py:515); anything else materialises the (B, N, E) tensor in the coordinates' dtype."""

    def __init__(self, module, coords):
        self.module, self.coords = module, coords
        self._val = None

    @property
    def shape(self):
        cs = self.coords.shape
        n = 1
        for v in cs[1:-1]:
            n *= v
        return torch.Size((cs[0], n // self.module.n_points, self.module.embedding_size))

    @property
    def dtype(self):
        return self.coords.dtype

    @property
    def device(self):
        return self.coords.device

    def _pe(self, tokens, out_dtype, device):
        from . import ops

        m, c = self.module, self.coords
        if isinstance(c, LazyPooled):
            fe = c.frontend(device, m.embedding_size, m.temperature, out_dtype)
            depth, poses, intr, clamp, resampled = c.inputs(device)
            if _dispatch(tokens):
                return torch.ops.v3d.frontend_pe(depth, poses, intr, tokens, ops.frontend_key(fe), clamp, c.dtype, resampled)
            return fe(depth, poses, intr, tokens=tokens, clamp=ops._clamp_arg(clamp), resampled=resampled, coord_dtype=c.dtype)
        x = c.reshape(c.shape[0], -1, 3)
        if _dispatch(tokens):
            return torch.ops.v3d.pe_sine3d(x, tokens, m.embedding_size, m.n_points, float(m.temperature), m.table_rows, out_dtype)
        return ops.pe_sine3d_impl(x, tokens, m.embedding_size, m.n_points, float(m.temperature), m.table_rows, out_dtype)

    def materialize(self) -> torch.Tensor:
        if self._val is None:
            from .video_utils import _device_for

            dev = self.coords.device if self.coords.device.type == "cuda" else _device_for()
            self._val = self._pe(None, self.dtype, dev)
        return self._val

    def add_to(self, tokens: torch.Tensor):
        """tokens + PE, fused; None when the combination has to go through torch's type promotion instead."""
        if not (isinstance(tokens, torch.Tensor) and tokens.is_cuda and tokens.dtype in _F and tuple(tokens.shape) == tuple(self.shape)):
            return None
        # A lazy depth handle that nobody cast (the HF Trainer only moves real tensors) keeps fp32 geometry and the PE is
        # rounded to the tokens' dtype; otherwise the reference's result dtype is promote(tokens, coords): fuse only when
        # that is the tokens' dtype.
        lazy_fp32 = isinstance(self.coords, LazyPooled) and self.dtype == torch.float32
        if not lazy_fp32 and torch.promote_types(tokens.dtype, self.dtype) != tokens.dtype:
            return None
        if not lazy_fp32 and self.dtype != tokens.dtype:
            return None  # e.g. fp32 tokens + bf16 PE: the reference rounds the PE to bf16 first
        return self._pe(tokens, tokens.dtype, tokens.device)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        if func in _add_funcs() and not kwargs and len(args) == 2:
            a, b = args
            pe, other = (a, b) if isinstance(a, LazyPE) else (b, a)
            if isinstance(pe, LazyPE) and isinstance(other, LazyTokenPool):
                out = other.with_pe(pe)
                if out is not None:
                    return out
            if isinstance(pe, LazyPE) and not isinstance(other, DeferredTensor):
                out = pe.add_to(other)
                if out is not None:
                    return out
        return func(*_unwrap(args), **_unwrap(kwargs or {}))

    def __add__(self, other):
        out = other.with_pe(self) if isinstance(other, LazyTokenPool) else self.add_to(other)
        return out if out is not None else self.materialize() + _unwrap(other)

    __radd__ = __add__


class LazyTokenPool(DeferredTensor):
    """Result of get_2dPool (llava/model/llava_arch.py:191-210) under the overlay: the pooled video tokens, not yet
    computed.  `+ LazyPE` (:515) keeps deferring; add_token_per_grid (:307-328) then runs pool + PE + newline as ONE
    kernel over the un-pooled tokens (v3d_pool_pe_grid) instead of three passes."""

    def __init__(self, tokens: torch.Tensor, stride: int, mode: str, pe: LazyPE | None = None):
        self.tokens, self.stride, self.mode, self.pe = tokens, int(stride), mode, pe
        self._val = None

    @property
    def out_grid(self) -> int:
        import math

        G = math.isqrt(self.tokens.shape[1])
        return -(-G // self.stride) if self.mode == "bilinear" else G // self.stride

    @property
    def shape(self):
        return torch.Size((self.tokens.shape[0], self.out_grid ** 2, self.tokens.shape[2]))

    @property
    def dtype(self):
        return self.tokens.dtype

    @property
    def device(self):
        return self.tokens.device

    @property
    def requires_grad(self):
        return self.tokens.requires_grad

    def detach(self):
        return LazyTokenPool(self.tokens.detach(), self.stride, self.mode, self.pe)

    def _pe_args(self):
        """(coords tensor (V, Go*Go, n, 3), module) of the deferred PE, or (None, None)."""
        if self.pe is None:
            return None, None
        c = self.pe.coords
        c = c.materialize() if isinstance(c, DeferredTensor) else c
        return c, self.pe.module

    def finish(self, image_newline=None, flatten=True) -> torch.Tensor:
        from .projector_tail import pool_pe_newline

        coords, module = self._pe_args()
        return pool_pe_newline(self.tokens, stride=self.stride, mode=self.mode, coords=coords, position_embedding=module,
                               image_newline=image_newline, flatten=flatten)

    def materialize(self) -> torch.Tensor:
        if self._val is None:
            self._val = self.finish()
        return self._val

    def with_pe(self, pe: LazyPE):
        """self + pe, still deferred -- or None when the PE cannot be fused into the pooling kernel."""
        ok = (self.pe is None and tuple(pe.shape) == tuple(self.shape) and self.tokens.is_cuda and self.dtype in _F
              and (pe.dtype == self.dtype or (isinstance(pe.coords, LazyPooled) and pe.dtype == torch.float32)))
        return LazyTokenPool(self.tokens, self.stride, self.mode, pe) if ok else None

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        if func in _add_funcs() and not kwargs and len(args) == 2:
            a, b = args
            pool, other = (a, b) if isinstance(a, LazyTokenPool) else (b, a)
            if isinstance(pool, LazyTokenPool) and isinstance(other, LazyPE):
                out = pool.with_pe(other)
                if out is not None:
                    return out
        return func(*_unwrap(args), **_unwrap(kwargs or {}))

    def __add__(self, other):
        if isinstance(other, LazyPE):
            out = self.with_pe(other)
            if out is not None:
                return out
        return self.materialize() + _unwrap(other)

    __radd__ = __add__

"""Deferred form of the object-proposal loop the reference inlines in prepare_inputs_labels_for_multimodal
(llava/model/llava_arch.py:351-378), so that the UNCHANGED loop ends in one kernel launch:

    world_coords = video_dict["world_coords"][0]                                   # a LazyWorldCoords scene handle
    for l in range(obj_num):
        min_xyz, max_xyz = box[:3] - box[3:] / 2, box[:3] + box[3:] / 2
        w = world_coords[:, :378, :378, :].reshape(-1, 27, 14, 27, 14, 3).transpose(2, 3).flatten(3, 4)   # PatchPixels
        m = torch.all((min_xyz <= w) & (w <= max_xyz), dim=-1)                     # BoxTest -> InBox
        m = m.sum(dim=3) >= int(14 * 14 * 0.5)                                     # Count -> ObjectMask (deferred)
        object_patch.append(m)
    ...
    image_features[0][object_patch[l].view(-1, 196)]                               # first .view(): ONE launch for all boxes

Every object below records one step of that chain; the first use of any ObjectMask's value evaluates all masks pending on
the handle with v3d_object_patch_masks_depth_bounds -- straight from depth, with the bounds the loop computed itself (in its
own dtype), so neither the (V,384,384,3) tensor nor the ~6 ATen launches per object happen.  Anything the chain does not
model falls back to real tensors (the handle materialises), so semantics never change.
"""
from __future__ import annotations

import torch

from . import cabi


def _is_tensor(x):
    return isinstance(x, torch.Tensor)


class _Deferred:
    """Base: unknown attributes / torch functions act on the real tensor."""

    def _real(self):
        raise NotImplementedError

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self._real(), name)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        name = getattr(func, "__name__", "")
        handled = cls._handle(name, args, kwargs)
        if handled is not NotImplemented:
            return handled
        real = lambda a: a._real() if isinstance(a, _Deferred) else a  # noqa: E731
        args = tuple(real(a) for a in args)
        kwargs = {k: real(v) for k, v in kwargs.items()}
        return func(*args, **kwargs)

    @classmethod
    def _handle(cls, name, args, kwargs):
        return NotImplemented


class PatchGrid(_Deferred):
    """world_coords[:, :n, :n, :] and the reshape / transpose / flatten that turn it into per-patch pixel lists."""

    def __init__(self, handle, n, ops=(), patch=None, grid=None):
        self._handle_obj, self._n, self._ops, self._patch, self._grid = handle, n, tuple(ops), patch, grid

    def _real(self):
        t = self._handle_obj.materialize()[:, : self._n, : self._n, :]
        for op, a in self._ops:
            t = getattr(t, op)(*a)
        return t

    def _next(self, op, args, **kw):
        return PatchGrid(self._handle_obj, self._n, self._ops + ((op, args),), kw.get("patch", self._patch), kw.get("grid", self._grid))

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
            shape = tuple(shape[0])
        if not self._ops and len(shape) == 6 and shape[0] == -1 and shape[5] == 3 and shape[1] == shape[3] and shape[2] == shape[4] \
                and shape[1] * shape[2] == self._n:
            return self._next("reshape", tuple(shape), patch=int(shape[2]), grid=int(shape[1]))
        return self._real().reshape(*shape)

    def transpose(self, a, b):
        if len(self._ops) == 1 and (a, b) in ((2, 3), (3, 2)):
            return self._next("transpose", (a, b))
        return self._real().transpose(a, b)

    def flatten(self, a=0, b=-1):
        if len(self._ops) == 2 and (a, b) == (3, 4):
            return PatchPixels(self._handle_obj, self._n, self._ops + (("flatten", (a, b)),), self._patch, self._grid)
        return self._real().flatten(a, b)


class PatchPixels(PatchGrid):
    """(V, G, G, p*p, 3): the pixels of every patch; comparisons against a 3-vector are recorded."""

    def __le__(self, other):  # w <= max_xyz
        return BoxTest(self, hi=other) if _is_tensor(other) and other.numel() == 3 else self._real() <= other

    def __ge__(self, other):  # w >= min_xyz  (reflected form of min_xyz <= w)
        return BoxTest(self, lo=other) if _is_tensor(other) and other.numel() == 3 else self._real() >= other

    @classmethod
    def _handle(cls, name, args, kwargs):
        # torch.Tensor.__le__(min_xyz, w) / torch.le(min_xyz, w): the tensor is on the left
        if name in ("__le__", "le") and len(args) == 2 and _is_tensor(args[0]) and isinstance(args[1], PatchPixels) and args[0].numel() == 3:
            return BoxTest(args[1], lo=args[0])
        if name in ("__ge__", "ge") and len(args) == 2 and _is_tensor(args[0]) and isinstance(args[1], PatchPixels) and args[0].numel() == 3:
            return BoxTest(args[1], hi=args[0])
        return NotImplemented


class BoxTest(_Deferred):
    """(lo <= w) and / or (w <= hi), element-wise over the last dimension."""

    def __init__(self, pixels, lo=None, hi=None, reduced=False, counted=False):
        self._pixels, self._lo, self._hi, self._reduced, self._counted = pixels, lo, hi, reduced, counted

    def _real(self):
        w = self._pixels._real()
        t = None
        if self._lo is not None:
            t = self._lo <= w
        if self._hi is not None:
            t = (w <= self._hi) if t is None else (t & (w <= self._hi))
        if self._reduced:
            t = torch.all(t, dim=-1)
        if self._counted:
            t = t.sum(dim=3)
        return t

    def __and__(self, other):
        if isinstance(other, BoxTest) and other._pixels is self._pixels and not (self._reduced or other._reduced) \
                and (self._lo is None) != (other._lo is None) and (self._hi is None) != (other._hi is None):
            return BoxTest(self._pixels, self._lo if self._lo is not None else other._lo, self._hi if self._hi is not None else other._hi)
        return self._real() & (other._real() if isinstance(other, _Deferred) else other)

    __rand__ = __and__

    def all(self, dim=None, **kw):
        if dim == -1 and not self._reduced and self._lo is not None and self._hi is not None and not kw:
            return BoxTest(self._pixels, self._lo, self._hi, reduced=True)
        return self._real().all(dim=dim, **kw)

    def sum(self, dim=None, **kw):
        if self._reduced and not self._counted and dim == 3 and not kw:
            return BoxTest(self._pixels, self._lo, self._hi, reduced=True, counted=True)
        return self._real().sum(dim=dim, **kw)

    def __ge__(self, k):
        if self._counted and isinstance(k, int):
            return ObjectMask(self._pixels, self._lo, self._hi, int(k))
        return self._real() >= k

    @classmethod
    def _handle(cls, name, args, kwargs):
        if name == "all" and len(args) >= 1 and isinstance(args[0], BoxTest):
            dim = kwargs.get("dim", args[1] if len(args) > 1 else None)
            extra = {k: v for k, v in kwargs.items() if k != "dim"}
            if not extra:
                return args[0].all(dim=dim)
        return NotImplemented


class ObjectMask(_Deferred):
    """One object's (V, G, G) patch mask, evaluated together with every other mask pending on the same scene handle."""

    def __init__(self, pixels, lo, hi, threshold):
        self._pixels, self._lo, self._hi, self._thr, self._val = pixels, lo, hi, threshold, None
        h = pixels._handle_obj
        h.__dict__.setdefault("_pending_object_masks", []).append(self)

    def _real(self):
        if self._val is None:
            evaluate_pending(self._pixels._handle_obj, self)
        return self._val

    def view(self, *shape):
        return self._real().view(*shape)


def evaluate_pending(handle, want: ObjectMask) -> None:
    """One v3d_object_patch_masks_depth_bounds launch for every pending mask of `handle` with want's patch geometry."""
    from .video_utils import _device_for

    pending = handle.__dict__.get("_pending_object_masks", [])
    key = (want._pixels._patch, want._pixels._grid, want._thr, want._pixels._n)
    group = [m for m in pending if m._val is None and (m._pixels._patch, m._pixels._grid, m._thr, m._pixels._n) == key]
    if want not in group:
        group.append(want)
    patch, grid, thr, _ = key
    lib = cabi.load()
    dev = want._lo.device if want._lo.is_cuda else _device_for(handle.depth)
    depth = handle._full_depth(dev).contiguous()  # the object kernel reads the frame layout
    poses, intr = handle.poses.to(dev, torch.float32).contiguous(), handle.intrinsics.to(dev, torch.float32).contiguous()
    V, H, W = depth.shape
    rm, cm = handle._maps(dev)
    bounds = torch.stack([torch.cat((m._lo.detach().reshape(3).to(dev, torch.float32), m._hi.detach().reshape(3).to(dev, torch.float32)))
                          for m in group]).contiguous()
    K = bounds.shape[0]
    masks = torch.zeros((K, V, grid, grid), dtype=torch.uint8, device=dev)
    lo_hi = None if handle.clamp is None else cabi.host_floats(list(handle.clamp[0]) + list(handle.clamp[1]))
    with torch.cuda.device(dev):
        cabi.check(lib.v3d_object_patch_masks_depth_bounds(depth.data_ptr(), cabi.depth_code(depth), poses.data_ptr(), intr.data_ptr(), V, H, W,
                                                           1000.0, rm.data_ptr(), cm.data_ptr(), lo_hi, cabi.dtype_code(handle.dtype), patch, grid,
                                                           bounds.data_ptr(), K, thr, masks.data_ptr(), cabi.stream_ptr(dev)),
                   "v3d_object_patch_masks_depth_bounds")
    masks = masks.bool()
    for i, m in enumerate(group):
        m._val = masks[i]
    handle.__dict__["_pending_object_masks"] = [m for m in pending if m._val is None]

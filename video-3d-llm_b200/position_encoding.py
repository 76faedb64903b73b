"""Mirror of llava/model/position_encoding.py on the CUDA library.

Same class names, constructor arguments, forward signature, output dtype/shape and (for the
MLP variant) state_dict keys as the reference, so `LlavaMetaModel.__init__`
(llava/model/llava_arch.py:48-63) can instantiate these instead.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import cabi
from . import ops  # noqa: F401  (registers torch.ops.v3d.*)
from .lazy import DeferredTensor, LazyPE, LazyPooled, _dispatch


#: where dim_t is evaluated.  The reference builds it on x.device (position_encoding.py:24), i.e. with torch's CUDA pow
#: for every real run; torch's fp32 pow is not correctly rounded and its CPU and CUDA results differ by one ulp for some
#: channels, which moves sin/cos of large coordinates by up to ~8e-6.  "device" (default, or V3D_DIM_T=device) follows
#: the reference on the GPU; "cpu" reproduces the reference run on the CPU (the golden vectors and the CPU oracle).
_dim_t_source = os.environ.get("V3D_DIM_T", "device")


def set_dim_t_source(source: str) -> str:
    """'device' or 'cpu'; returns the previous setting."""
    global _dim_t_source
    if source not in ("device", "cpu"):
        raise ValueError("dim_t source must be 'device' or 'cpu'")
    prev, _dim_t_source = _dim_t_source, source
    return prev


def sine3d_dim_t(num_feats: int, temperature: float = 10000, device="cpu") -> torch.Tensor:
    """llava/model/position_encoding.py:24-25, evaluated by torch exactly as the reference writes it."""
    dim_t = torch.arange(num_feats, dtype=torch.float32, device=device)
    return temperature ** (2 * (dim_t // 2) / num_feats)


class Sine3DTables:
    """dim_t and the lookup table of one (E, n_points, temperature, dtype) on one device."""

    def __init__(self, embedding_size: int, n_points: int, temperature: float, dtype: torch.dtype, device, rows: int):
        lib = cabi.load()
        self.E, self.n_points, self.rows, self.dtype = embedding_size, n_points, rows, dtype
        self.nf = embedding_size // (3 * n_points)
        self.dim_t = sine3d_dim_t(self.nf, temperature, device if _dim_t_source == "device" else "cpu").to(device)
        nbytes = lib.v3d_pe_table_bytes(embedding_size, n_points, rows, cabi.dtype_code(dtype))
        if nbytes == 0:
            raise NotImplementedError(f"no PE table for E={embedding_size}, n_points={n_points} (needs E // (3*n_points) >= 8)")
        self.table = torch.empty(nbytes, dtype=torch.uint8, device=device)
        with torch.cuda.device(device):
            cabi.check(lib.v3d_pe_table_build(self.dim_t.data_ptr(), embedding_size, n_points, rows, cabi.dtype_code(dtype),
                                              self.table.data_ptr(), cabi.stream_ptr(device)), "v3d_pe_table_build")


_tables: dict = {}


def get_tables(embedding_size, n_points, temperature, dtype, device, rows) -> Sine3DTables:
    key = (embedding_size, n_points, float(temperature), dtype, str(torch.device(device)), rows, _dim_t_source)
    if key not in _tables:
        _tables[key] = Sine3DTables(embedding_size, n_points, temperature, dtype, device, rows)
    return _tables[key]


class PositionEmbeddingSine3D(nn.Module):
    """llava/model/position_encoding.py:5-49.  forward(x): (B,N,3) or (B,N,n_points,3) ->
    (B,N,embedding_size) in x.dtype; no parameters.  `add_to` (same shape/dtype as the output)
    fuses the add of llava_arch.py:515 into the same kernel."""

    #: integer coordinates below this bound are served from the lookup table
    table_rows = 512
    #: True (set by the overlay launcher): forward returns a deferred lazy.LazyPE so that the `tokens + PE` of
    #: llava_arch.py:515 becomes one fused launch; a deferred coordinate input always yields a deferred PE
    defer_add = False

    def __init__(self, embedding_size, temperature=10000, n_points=1):
        super().__init__()
        self.embedding_size = embedding_size
        self.temperature = temperature
        self.n_points = n_points

    def forward(self, x, add_to=None):
        if isinstance(x, LazyPooled) or (self.defer_add and add_to is None and x.is_cuda):
            pe = LazyPE(self, x)
            return pe if add_to is None else pe + add_to
        cabi.require_cuda(x, "PositionEmbeddingSine3D input")
        if self.n_points > 1:
            x = x.flatten(1, 2)
        out_dtype = x.dtype
        if add_to is not None and (add_to.dtype != out_dtype or tuple(add_to.shape) != (x.shape[0], x.shape[1] // self.n_points, self.embedding_size)):
            raise ValueError("add_to must have the output's shape and dtype")
        if _dispatch(add_to):
            return torch.ops.v3d.pe_sine3d(x, add_to, self.embedding_size, self.n_points, float(self.temperature), self.table_rows, out_dtype)
        return ops.pe_sine3d_impl(x, add_to, self.embedding_size, self.n_points, float(self.temperature), self.table_rows, out_dtype)


class PositionEmbeddingMLP(nn.Module):
    """llava/model/position_encoding.py:52-84: Linear(3,512) -> LayerNorm -> ReLU -> Linear(512,E).
    Parameters live under mlp.{0,1,3}.{weight,bias} like the reference (train_3d.py:1818-1821 makes
    them trainable).  A bf16 module runs the tcgen05 kernel (hidden activations never reach HBM) in
    training and inference alike: torch.ops.v3d.pe_mlp carries the backward (ops.py)."""

    def __init__(self, embedding_size, temperature=10000, n_points=1):
        super().__init__()
        self.embedding_size = embedding_size
        self.temperature = temperature
        self.n_points = n_points
        self.hidden_size = 512
        self.mlp = nn.Sequential(
            nn.Linear(3, self.hidden_size),
            nn.LayerNorm(self.hidden_size),
            nn.ReLU(),
            nn.Linear(self.hidden_size, self.embedding_size),
        )

    def forward(self, x, add_to=None):
        if self.n_points > 1:
            # the reference's n_points > 1 branch dereferences an undefined name (position_encoding.py:79)
            raise NotImplementedError("PositionEmbeddingMLP supports n_points == 1 only, like the reference")
        if isinstance(x, DeferredTensor):
            x = x.materialize()
        cabi.require_cuda(x, "PositionEmbeddingMLP input")
        lin1, ln, lin2 = self.mlp[0], self.mlp[1], self.mlp[3]
        # The tcgen05 kernel multiplies with W2 in bf16 and emits bf16: it serves bf16 modules (the recipe trains and
        # evaluates in bf16).  Any other dtype runs the reference's own operator sequence on the device, so an fp32 / fp16
        # model keeps its precision; train and eval go through the same path for a given dtype.
        if lin2.weight.dtype != torch.bfloat16 or x.dtype != torch.bfloat16:
            pos = self.mlp(x)
            return pos if add_to is None else add_to + pos
        B, N, _ = x.shape
        tok = None
        if add_to is not None:
            if add_to.dtype != torch.bfloat16 or tuple(add_to.shape) != (B, N, self.embedding_size):
                pos = self.forward(x)
                return add_to + pos
            tok = add_to.reshape(B * N, self.embedding_size)
        out = torch.ops.v3d.pe_mlp(x.detach().reshape(B * N, 3), lin1.weight, lin1.bias, ln.weight, ln.bias, float(ln.eps), lin2.weight,
                                   lin2.bias, tok)
        return out.view(B, N, self.embedding_size)

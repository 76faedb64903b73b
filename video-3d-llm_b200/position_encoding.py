"""Mirror of llava/model/position_encoding.py on the CUDA library.

Same class names, constructor arguments, forward signature, output dtype/shape and (for the
MLP variant) state_dict keys as the reference, so `LlavaMetaModel.__init__`
(llava/model/llava_arch.py:48-63) can instantiate these instead.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import cabi


def sine3d_dim_t(num_feats: int, temperature: float = 10000) -> torch.Tensor:
    """llava/model/position_encoding.py:24-25, evaluated by torch on the CPU exactly as the reference
    writes it (torch's fp32 pow is not correctly rounded, so this is the bit-exact source)."""
    dim_t = torch.arange(num_feats, dtype=torch.float32)
    return temperature ** (2 * (dim_t // 2) / num_feats)


class Sine3DTables:
    """dim_t and the lookup table of one (E, n_points, temperature, dtype) on one device."""

    def __init__(self, embedding_size: int, n_points: int, temperature: float, dtype: torch.dtype, device, rows: int):
        lib = cabi.load()
        self.E, self.n_points, self.rows, self.dtype = embedding_size, n_points, rows, dtype
        self.nf = embedding_size // (3 * n_points)
        self.dim_t = sine3d_dim_t(self.nf, temperature).to(device)
        nbytes = lib.v3d_pe_table_bytes(embedding_size, n_points, rows, cabi.dtype_code(dtype))
        if nbytes == 0:
            raise NotImplementedError(f"no PE table for E={embedding_size}, n_points={n_points} (needs E // (3*n_points) >= 8)")
        self.table = torch.empty(nbytes, dtype=torch.uint8, device=device)
        with torch.cuda.device(device):
            cabi.check(lib.v3d_pe_table_build(self.dim_t.data_ptr(), embedding_size, n_points, rows, cabi.dtype_code(dtype),
                                              self.table.data_ptr(), cabi.stream_ptr(device)), "v3d_pe_table_build")


_tables: dict = {}


def get_tables(embedding_size, n_points, temperature, dtype, device, rows) -> Sine3DTables:
    key = (embedding_size, n_points, float(temperature), dtype, str(torch.device(device)), rows)
    if key not in _tables:
        _tables[key] = Sine3DTables(embedding_size, n_points, temperature, dtype, device, rows)
    return _tables[key]


class PositionEmbeddingSine3D(nn.Module):
    """llava/model/position_encoding.py:5-49.  forward(x): (B,N,3) or (B,N,n_points,3) ->
    (B,N,embedding_size) in x.dtype; no parameters.  `add_to` (same shape/dtype as the output)
    fuses the add of llava_arch.py:515 into the same kernel."""

    #: integer coordinates below this bound are served from the lookup table
    table_rows = 512

    def __init__(self, embedding_size, temperature=10000, n_points=1):
        super().__init__()
        self.embedding_size = embedding_size
        self.temperature = temperature
        self.n_points = n_points

    def forward(self, x, add_to=None):
        cabi.require_cuda(x, "PositionEmbeddingSine3D input")
        lib = cabi.load()
        if self.n_points > 1:
            x = x.flatten(1, 2)
        B, N, _ = x.shape
        n_tok = N // self.n_points
        out_dtype = x.dtype
        coords = x.detach().to(torch.float32).contiguous()
        tabs = get_tables(self.embedding_size, self.n_points, self.temperature, out_dtype, x.device, self.table_rows)
        out = torch.empty((B, n_tok, self.embedding_size), dtype=out_dtype, device=x.device)
        if add_to is not None:
            if add_to.shape != out.shape or add_to.dtype != out_dtype:
                raise ValueError("add_to must have the output's shape and dtype")
            add_to = add_to.contiguous()
        with torch.cuda.device(x.device):
            cabi.check(lib.v3d_pe_sine3d(coords.data_ptr(), B * n_tok, self.n_points, self.embedding_size,
                                         tabs.dim_t.data_ptr(), tabs.table.data_ptr(), tabs.rows, cabi.dtype_code(out_dtype),
                                         cabi.ptr(add_to), out.data_ptr(), cabi.stream_ptr(x.device)), "v3d_pe_sine3d")
        return out


class PositionEmbeddingMLP(nn.Module):
    """llava/model/position_encoding.py:52-84: Linear(3,512) -> LayerNorm -> ReLU -> Linear(512,E).
    Parameters live under mlp.{0,1,3}.{weight,bias} like the reference (train_3d.py:1818-1821 makes
    them trainable).  Inference runs the tcgen05 kernel (hidden activations never reach HBM);
    when gradients are needed the same math runs through torch autograd."""

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
        cabi.require_cuda(x, "PositionEmbeddingMLP input")
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if needs_grad:
            pos = self.mlp(x)
            return pos if add_to is None else add_to + pos
        lib = cabi.load()
        B, N, _ = x.shape
        lin1, ln, lin2 = self.mlp[0], self.mlp[1], self.mlp[3]
        coords = x.detach().to(torch.float32).contiguous()
        f32 = lambda t: t.detach().to(torch.float32).contiguous()
        w2 = lin2.weight.detach().to(torch.bfloat16).contiguous()
        out = torch.empty((B, N, self.embedding_size), dtype=torch.bfloat16, device=x.device)
        if add_to is not None:
            add_to = add_to.to(torch.bfloat16).contiguous()
        w1, b1, g, be, b2 = f32(lin1.weight), f32(lin1.bias), f32(ln.weight), f32(ln.bias), f32(lin2.bias)
        with torch.cuda.device(x.device):
            cabi.check(lib.v3d_pe_mlp(coords.data_ptr(), B * N, self.embedding_size, w1.data_ptr(), b1.data_ptr(), g.data_ptr(),
                                      be.data_ptr(), float(ln.eps), w2.data_ptr(), b2.data_ptr(), cabi.ptr(add_to),
                                      out.data_ptr(), cabi.stream_ptr(x.device)), "v3d_pe_mlp")
        return out.to(x.dtype)

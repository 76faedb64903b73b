"""torch.library registration of the hot-path entry points (SURVEY.md section 8b "Autograd / compile").

The C ABI is bound with ctypes (cabi.py); these thin `torch.library.custom_op` wrappers make the three operators
that sit inside the model's forward opaque, traceable nodes for `torch.compile` (the recipe passes
`--torch_compile True`, scripts/3d/train/train_multi.sh:72-76) and give them autograd formulas:

  v3d::frontend_pe   depth/pose/intrinsics (+ tokens) -> tokens + sin-3D PE   (video_utils.py:230,296-308 ->
                     llava_arch.py:396-420 -> position_encoding.py:17-49 -> llava_arch.py:515); the PE has no
                     parameters and its input is detached (:272,515): d out / d tokens = identity
  v3d::frontend_coords   the same kernel with the PE switched off: pooled + discretised coordinates
  v3d::pe_sine3d     PositionEmbeddingSine3D.forward on coordinates (+ tokens): same backward
  v3d::pool_pe_grid  get_2dPool -> + PE -> add_token_per_grid in one kernel (llava_arch.py:191-210, 515, 307-328); backward
                     = the reference's operator sequence differentiated by torch (the projector and the newline
                     parameter are trained, so the tokens and image_newline need gradients)
  v3d::pe_mlp        PositionEmbeddingMLP.forward (+ tokens) on tcgen05; backward = the reference's own operator
                     sequence differentiated by torch (the module's parameters are trainable, train_3d.py:1818-1821)

Fake (meta) implementations only describe output shapes/dtypes, so tracing never touches the library.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from . import cabi

_frontends: dict = {}
_frontend_keys: dict = {}


def frontend_key(fe) -> int:
    """Register a FrontEnd (configuration + cached maps/tables) and return the integer the ops refer to it by."""
    k = _frontend_keys.get(id(fe))
    if k is None:
        k = len(_frontends)
        _frontends[k] = fe
        _frontend_keys[id(fe)] = k
    return k


def _clamp_arg(clamp: List[float]):
    return None if len(clamp) == 0 else (clamp[:3], clamp[3:])


# ---- fused front end ---------------------------------------------------------------------------------------------
@torch.library.custom_op("v3d::frontend_pe", mutates_args=())
def frontend_pe(depth: torch.Tensor, poses: torch.Tensor, intrinsics: torch.Tensor, tokens: Optional[torch.Tensor], fe: int,
                clamp: List[float], coord_dtype: torch.dtype, resampled: bool) -> torch.Tensor:
    f = _frontends[fe]
    return f(depth, poses, intrinsics, tokens=tokens, clamp=_clamp_arg(clamp), resampled=resampled, coord_dtype=coord_dtype)


@frontend_pe.register_fake
def _(depth, poses, intrinsics, tokens, fe, clamp, coord_dtype, resampled):
    f = _frontends[fe]
    return depth.new_empty((depth.shape[0], f.tokens_per_frame, f.cfg.embedding_size), dtype=f.cfg.dtype)


def _frontend_pe_setup(ctx, inputs, output):
    ctx.has_tokens = inputs[3] is not None


def _frontend_pe_backward(ctx, grad):
    return None, None, None, (grad if ctx.has_tokens else None), None, None, None, None


torch.library.register_autograd("v3d::frontend_pe", _frontend_pe_backward, setup_context=_frontend_pe_setup)


@torch.library.custom_op("v3d::frontend_coords", mutates_args=())
def frontend_coords(depth: torch.Tensor, poses: torch.Tensor, intrinsics: torch.Tensor, fe: int, clamp: List[float],
                    coord_dtype: torch.dtype, resampled: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """(pooled, pe_coords), both fp32 (V,G,G[,n],3): what the reference's average_coordinate_in_patch & co. and
    discrete_coords return, values already rounded to coord_dtype."""
    f = _frontends[fe]
    out = f(depth, poses, intrinsics, clamp=_clamp_arg(clamp), want_coords=True, resampled=resampled, coord_dtype=coord_dtype, pe=False)
    return out[1], out[2]


@frontend_coords.register_fake
def _(depth, poses, intrinsics, fe, clamp, coord_dtype, resampled):
    f = _frontends[fe]
    G = f.cfg.grid
    shape = (depth.shape[0], G, G, 3) if f.n_points == 1 else (depth.shape[0], G, G, f.n_points, 3)
    return depth.new_empty(shape, dtype=torch.float32), depth.new_empty(shape, dtype=torch.float32)


# ---- stand-alone sin-3D PE ----------------------------------------------------------------------------------------
def pe_sine3d_impl(coords: torch.Tensor, tokens: Optional[torch.Tensor], embedding_size: int, n_points: int, temperature: float,
                   table_rows: int, out_dtype: torch.dtype) -> torch.Tensor:
    """coords (B, N*n_points, 3) any float dtype -> (B, N, E) in out_dtype (+ tokens)."""
    from .position_encoding import get_tables

    lib = cabi.load()
    B, N, _ = coords.shape
    n_tok = N // n_points
    c = coords.detach().to(torch.float32).contiguous()
    tabs = get_tables(embedding_size, n_points, temperature, out_dtype, coords.device, table_rows)
    out = torch.empty((B, n_tok, embedding_size), dtype=out_dtype, device=coords.device)
    if tokens is not None:
        if tokens.shape != out.shape or tokens.dtype != out_dtype:
            raise ValueError(f"tokens must be {tuple(out.shape)} {out_dtype}, got {tuple(tokens.shape)} {tokens.dtype}")
        tokens = tokens.contiguous()
    with torch.cuda.device(coords.device):
        cabi.check(lib.v3d_pe_sine3d(c.data_ptr(), B * n_tok, n_points, embedding_size, tabs.dim_t.data_ptr(), tabs.table.data_ptr(),
                                     tabs.rows, cabi.dtype_code(out_dtype), cabi.ptr(tokens), out.data_ptr(),
                                     cabi.stream_ptr(coords.device)), "v3d_pe_sine3d")
    return out


pe_sine3d = torch.library.custom_op("v3d::pe_sine3d", mutates_args=())(pe_sine3d_impl)


@pe_sine3d.register_fake
def _(coords, tokens, embedding_size, n_points, temperature, table_rows, out_dtype):
    return coords.new_empty((coords.shape[0], coords.shape[1] // n_points, embedding_size), dtype=out_dtype)


def _pe_sine3d_setup(ctx, inputs, output):
    ctx.has_tokens = inputs[1] is not None


def _pe_sine3d_backward(ctx, grad):
    return None, (grad if ctx.has_tokens else None), None, None, None, None, None


torch.library.register_autograd("v3d::pe_sine3d", _pe_sine3d_backward, setup_context=_pe_sine3d_setup)


# ---- learned-projection PE ----------------------------------------------------------------------------------------
@torch.library.custom_op("v3d::pe_mlp", mutates_args=())
def pe_mlp(coords: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float,
           w2: torch.Tensor, b2: torch.Tensor, tokens: Optional[torch.Tensor]) -> torch.Tensor:
    """coords (M,3); w1 (512,3), b1, gamma, beta (512); w2 (E,512); b2 (E); tokens None or (M,E) bf16 -> (M,E) bf16."""
    lib = cabi.load()
    M, E = coords.shape[0], w2.shape[0]
    f32 = lambda t: t.detach().to(torch.float32).contiguous()  # noqa: E731
    c, w1f, b1f, gf, bf, b2f = f32(coords), f32(w1), f32(b1), f32(gamma), f32(beta), f32(b2)
    w2h = w2.detach().to(torch.bfloat16).contiguous()
    out = torch.empty((M, E), dtype=torch.bfloat16, device=coords.device)
    if tokens is not None:
        if tokens.shape != out.shape or tokens.dtype != torch.bfloat16:
            raise ValueError(f"tokens must be {tuple(out.shape)} bfloat16, got {tuple(tokens.shape)} {tokens.dtype}")
        tokens = tokens.contiguous()
    with torch.cuda.device(coords.device):
        cabi.check(lib.v3d_pe_mlp(c.data_ptr(), M, E, w1f.data_ptr(), b1f.data_ptr(), gf.data_ptr(), bf.data_ptr(), float(eps),
                                  w2h.data_ptr(), b2f.data_ptr(), cabi.ptr(tokens), out.data_ptr(), cabi.stream_ptr(coords.device)),
                   "v3d_pe_mlp")
    return out


@pe_mlp.register_fake
def _(coords, w1, b1, gamma, beta, eps, w2, b2, tokens):
    return coords.new_empty((coords.shape[0], w2.shape[0]), dtype=torch.bfloat16)


def _pe_mlp_setup(ctx, inputs, output):
    coords, w1, b1, gamma, beta, eps, w2, b2, tokens = inputs
    ctx.save_for_backward(coords, w1, b1, gamma, beta, w2, b2)
    ctx.eps = eps
    ctx.has_tokens = tokens is not None


def _pe_mlp_backward(ctx, grad):
    """Backward of position_encoding.py:60-65 by recomputing the hidden activations with torch (fp32 statistics like
    nn.LayerNorm under autocast-free bf16 training) and differentiating the reference's operator sequence."""
    coords, w1, b1, gamma, beta, w2, b2 = ctx.saved_tensors
    params = [w1, b1, gamma, beta, w2, b2]
    need = ctx.needs_input_grad
    with torch.enable_grad():
        leaves = [p.detach().requires_grad_(True) for p in params]
        lw1, lb1, lg, lbt, lw2, lb2 = leaves
        h = torch.nn.functional.linear(coords.detach().to(lw1.dtype), lw1, lb1)
        h = torch.nn.functional.layer_norm(h, (h.shape[-1],), lg, lbt, ctx.eps)
        pos = torch.nn.functional.linear(torch.relu(h), lw2, lb2)
        grads = torch.autograd.grad(pos, leaves, grad.to(pos.dtype), allow_unused=True)
    gw1, gb1, gg, gbt, gw2, gb2 = grads
    pick = lambda g, i: g if need[i] else None  # noqa: E731
    return (None, pick(gw1, 1), pick(gb1, 2), pick(gg, 3), pick(gbt, 4), None, pick(gw2, 6), pick(gb2, 7),
            grad if (ctx.has_tokens and need[8]) else None)


torch.library.register_autograd("v3d::pe_mlp", _pe_mlp_backward, setup_context=_pe_mlp_setup)


# ---- projector tail: token pooling (+ PE) (+ newline) ----------------------------------------------------------------
_POOL_NAMES = {cabi.TOKEN_POOL_BILINEAR: "bilinear", cabi.TOKEN_POOL_AVERAGE: "average", cabi.TOKEN_POOL_MAX: "max", cabi.TOKEN_POOL_NONE: "none"}


@torch.library.custom_op("v3d::pool_pe_grid", mutates_args=())
def pool_pe_grid(tokens: torch.Tensor, stride: int, mode: int, coords: Optional[torch.Tensor], n_points: int, temperature: float,
                 table_rows: int, newline: Optional[torch.Tensor]) -> torch.Tensor:
    """tokens (V, G*G, E) -> (V, Go, Go + (newline is not None), E); coords None or (V, Go*Go, n_points, 3) fp32."""
    from .position_encoding import get_tables

    lib = cabi.load()
    V, N, E = tokens.shape
    G = int(round(N ** 0.5))
    Go = lib.v3d_token_pool_out_grid(G, int(stride), int(mode))
    dt = tokens.dtype
    x = tokens.contiguous()
    d = cabi.PoolDesc()
    d.tokens, d.n_frames, d.src_grid, d.E, d.dtype = x.data_ptr(), V, G, E, cabi.dtype_code(dt)
    d.pool_mode, d.stride = int(mode), int(stride)
    keep = [x]
    if coords is not None:
        c = coords.detach().to(torch.float32).contiguous()
        tabs = get_tables(E, n_points, temperature, dt, x.device, table_rows)
        d.coords, d.n_points, d.dim_t, d.table, d.table_rows = c.data_ptr(), n_points, tabs.dim_t.data_ptr(), tabs.table.data_ptr(), tabs.rows
        keep += [c, tabs]
    nl = 0
    if newline is not None:
        nlv = newline.detach().to(device=x.device, dtype=dt).contiguous()
        d.newline, nl = nlv.data_ptr(), 1
        keep.append(nlv)
    out = torch.empty((V, Go, Go + nl, E), dtype=dt, device=x.device)
    d.out = out.data_ptr()
    with torch.cuda.device(x.device):
        cabi.check(lib.v3d_pool_pe_grid(d, cabi.stream_ptr(x.device)), "v3d_pool_pe_grid")
    return out


@pool_pe_grid.register_fake
def _(tokens, stride, mode, coords, n_points, temperature, table_rows, newline):
    import math

    V, N, E = tokens.shape
    G = math.isqrt(N)
    Go = G if mode == cabi.TOKEN_POOL_NONE else (-(-G // stride) if mode == cabi.TOKEN_POOL_BILINEAR else G // stride)
    return tokens.new_empty((V, Go, Go + (1 if newline is not None else 0), E))


def _pool_setup(ctx, inputs, output):
    tokens, stride, mode, coords, n_points, temperature, table_rows, newline = inputs
    ctx.stride, ctx.mode, ctx.has_newline = stride, mode, newline is not None
    ctx.tok_shape, ctx.tok_dtype = tokens.shape, tokens.dtype
    ctx.newline_dtype = newline.dtype if newline is not None else None
    ctx.save_for_backward(tokens if mode == cabi.TOKEN_POOL_MAX else None)  # only max pooling's VJP depends on the values


def _pool_backward(ctx, grad):
    """VJP of llava_arch.py:191-210 + :307-328 by differentiating the reference's own operator sequence with torch (the
    pooling is linear except for `max`, so a zero input serves as the linearisation point)."""
    import math

    (saved,) = ctx.saved_tensors
    V, N, E = ctx.tok_shape
    G = math.isqrt(N)
    g_nl = None
    if ctx.has_newline:
        g_nl = grad[:, :, -1, :].sum(dim=(0, 1)).to(ctx.newline_dtype) if ctx.needs_input_grad[7] else None
        grad = grad[:, :, :-1, :]
    g_tok = None
    if ctx.needs_input_grad[0]:
        Go = grad.shape[1]
        gy = grad.permute(0, 3, 1, 2)  # (V, E, Go, Go) like the reference's channel-first pooling
        if ctx.mode == cabi.TOKEN_POOL_NONE:
            g_tok = grad.reshape(V, N, E)
        else:
            with torch.enable_grad():
                x = (saved.detach() if saved is not None else torch.zeros((V, N, E), dtype=ctx.tok_dtype, device=grad.device)).requires_grad_(True)
                xc = x.view(V, G, G, E).permute(0, 3, 1, 2)
                if ctx.mode == cabi.TOKEN_POOL_AVERAGE:
                    y = torch.nn.functional.avg_pool2d(xc, ctx.stride)
                elif ctx.mode == cabi.TOKEN_POOL_MAX:
                    y = torch.nn.functional.max_pool2d(xc, ctx.stride)
                else:
                    y = torch.nn.functional.interpolate(xc, size=[Go, Go], mode="bilinear")
                (g_tok,) = torch.autograd.grad(y, x, gy.to(y.dtype))
    return g_tok, None, None, None, None, None, None, g_nl


torch.library.register_autograd("v3d::pool_pe_grid", _pool_backward, setup_context=_pool_setup)

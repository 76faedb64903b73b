"""Mirror of the projector tail of LlavaMetaForCausalLM (llava/model/llava_arch.py) on the CUDA library
(SURVEY.md section 8 f2): what happens to the projected tokens of a video between `mm_projector` and the language
model --

    get_2dPool(image_feature, stride=2)          :191-210   (V,27*27,E) -> (V,14*14,E)
    image_feat + world_position_embedding(c)     :506-517
    add_token_per_grid(image_feature)            :307-328   -> (V*14*15, E)

`get_2dPool` and `add_token_per_grid` keep the reference's names, arguments and results as functions of explicit
inputs (the reference reads config/model attributes from `self`); `pool_pe_newline` is the three steps in one
kernel launch.  The projector's two Linear layers stay with cuBLAS: they are plain library GEMMs.
"""
from __future__ import annotations

import math

import torch

from . import cabi
from . import ops  # noqa: F401  (registers torch.ops.v3d.*)
from .position_encoding import PositionEmbeddingSine3D

_MODES = {"bilinear": cabi.TOKEN_POOL_BILINEAR, "average": cabi.TOKEN_POOL_AVERAGE, "max": cabi.TOKEN_POOL_MAX, "none": cabi.TOKEN_POOL_NONE}


def pool_pe_newline(image_feature: torch.Tensor, *, stride: int = 2, mode: str = "bilinear", coords: torch.Tensor | None = None,
                    position_embedding: PositionEmbeddingSine3D | None = None, image_newline: torch.Tensor | None = None,
                    flatten: bool = True) -> torch.Tensor:
    """image_feature (V, G*G, E) on a CUDA device -> pooled (+ PE) (+ newline column).

    coords: what the PE is evaluated on, (V, Go*Go, 3) or (V, Go*Go, n_points, 3) (e.g. the discretised pooled
    coordinates); requires position_embedding.  image_newline (E,): appended after every grid row.
    Returns (V*Go*(Go+1), E) like add_token_per_grid when a newline is given and flatten is set, else (V, Go*Go', E)."""
    cabi.require_cuda(image_feature, "image_feature")
    if mode not in _MODES:
        raise ValueError(f"Unexpected mm_spatial_pool_mode: {mode}")
    lib = cabi.load()
    V, N, E = image_feature.shape
    G = int(math.isqrt(N))
    if G * G != N:
        raise ValueError(f"{N} tokens per frame do not form a square grid")
    Go = lib.v3d_token_pool_out_grid(G, int(stride), _MODES[mode])
    if Go <= 0:
        raise ValueError(f"bad stride {stride} for a {G}x{G} grid")
    c, n_points, temperature, rows = None, 1, 10000.0, 0
    if coords is not None:
        if position_embedding is None:
            raise ValueError("coords need the position_embedding module they are encoded with")
        if position_embedding.embedding_size != E:
            raise ValueError("embedding size of the PE module and of the tokens differ")
        n_points, temperature, rows = position_embedding.n_points, float(position_embedding.temperature), position_embedding.table_rows
        c = coords.detach().to(torch.float32).reshape(V, Go * Go, n_points, 3)
    if image_newline is not None and image_newline.shape != (E,):
        raise ValueError("image_newline must have shape (E,)")
    out = torch.ops.v3d.pool_pe_grid(image_feature, int(stride), _MODES[mode], c, n_points, temperature, rows, image_newline)
    nl = 0 if image_newline is None else 1
    if nl and flatten:
        return out.view(V * Go * (Go + 1), E)
    return out.view(V, Go * (Go + nl), E)


def get_2dPool(image_feature: torch.Tensor, stride: int = 2, mm_spatial_pool_mode: str = "bilinear") -> torch.Tensor:
    """llava_arch.py:191-210: (V, G*G, E) -> (V, Go*Go, E)."""
    return pool_pe_newline(image_feature, stride=stride, mode=mm_spatial_pool_mode)


def add_token_per_grid(image_feature: torch.Tensor, image_newline: torch.Tensor) -> torch.Tensor:
    """llava_arch.py:307-328 (add_faster_video off): (V, G*G, E) -> (V*G*(G+1), E)."""
    return pool_pe_newline(image_feature, mode="none", image_newline=image_newline)

"""Mirror of the projector tail of LlavaMetaForCausalLM (llava/model/llava_arch.py) on the CUDA library
(SURVEY.md section 8 f2): what happens to the projected tokens of a video between `mm_projector` and the language
model --

    get_2dPool(image_feature, stride=2)          :191-210   (V,27*27,E) -> (V,14*14,E)
    image_feat + world_position_embedding(c)     :506-517
    add_token_per_grid(image_feature)            :307-328   -> (V*14*15, E)

`get_2dPool` and `add_token_per_grid` keep the reference's names, arguments and results as functions of explicit
inputs (the reference reads config/model attributes from `self`); `pool_pe_newline` is the three steps in one
kernel launch.  The projector's two Linear layers stay with cuBLAS: they are plain library GEMMs -- but
`project_pool_pe_newline` runs the last one on the POOLED tokens (the bilinear / average pool is linear with weights
that sum to one, so it commutes with Linear + bias): 3.7x fewer GEMM rows for the 27x27 -> 14x14 recipe.
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


def project_pool_pe_newline(vision_features: torch.Tensor, mm_projector: torch.nn.Module, *, stride: int = 2, mode: str = "bilinear",
                            coords: torch.Tensor | None = None, position_embedding: PositionEmbeddingSine3D | None = None,
                            image_newline: torch.Tensor | None = None, flatten: bool = True) -> torch.Tensor:
    """mm_projector -> get_2dPool -> + PE -> add_token_per_grid (llava/model/multimodal_projector/builder.py:41-48,
    llava_arch.py:191-210, 515, 307-328) with the pool moved in front of the projector's last Linear.

    `get_2dPool(Linear(h)) == Linear(get_2dPool(h))` in real arithmetic for the bilinear and average modes (every output
    token is a convex combination of input tokens, and the bias passes through weights that sum to one), so the last
    Linear -- for `mlp2x_gelu` 3584 x 3584 per token, by far the most expensive step between the vision tower and the
    language model -- runs on Go*Go = 196 instead of G*G = 729 tokens per frame.  What changes is the rounding: the
    reference rounds the Linear's 729 outputs to the token dtype and pools them, this rounds the pooled hidden
    activations and the Linear's 196 outputs; the two results differ by a few ulps of the token dtype (tests/test_gpu_tail.py
    states the bound), which is why this is an opt-in entry point for a maintainer who edits llava_arch.py:469-536 and
    not what the overlay launcher installs.  The GEMMs stay library GEMMs (cuBLAS through torch).

    vision_features: (V, G*G, C) vision-tower output.  mm_projector: nn.Linear, or an nn.Sequential whose last module is an
    nn.Linear (`linear`, `mlpNx_gelu` of the reference's builder); anything else, and mode 'max', runs the reference order."""
    cabi.require_cuda(vision_features, "vision_features")
    last = mm_projector if isinstance(mm_projector, torch.nn.Linear) else None
    head = None
    if isinstance(mm_projector, torch.nn.Sequential) and len(mm_projector) > 0 and isinstance(mm_projector[-1], torch.nn.Linear):
        last, head = mm_projector[-1], mm_projector[:-1]
    if last is None or mode not in ("bilinear", "average"):
        return pool_pe_newline(mm_projector(vision_features), stride=stride, mode=mode, coords=coords, position_embedding=position_embedding,
                               image_newline=image_newline, flatten=flatten)
    h = vision_features if head is None else head(vision_features)
    pooled = pool_pe_newline(h.contiguous(), stride=stride, mode=mode)  # (V, Go*Go, hidden): k_pool_pe_grid, pool only
    y = last(pooled)
    return pool_pe_newline(y, mode="none", coords=coords, position_embedding=position_embedding, image_newline=image_newline, flatten=flatten)

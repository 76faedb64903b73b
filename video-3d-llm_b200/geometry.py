"""Host-side geometry of the front end: the nearest-neighbour resize + centre crop of
llava/video_utils.py:291-308 expressed as two index maps, and the patch grid.

cv2.INTER_NEAREST reads source index min(floor(dst * (1 / (dst_size / src_size))), src_size-1)
in double precision; for ScanNet (480x640 -> 384x512 -> crop 384 at left=64) that is
row_map[r] = floor(1.25 r), col_map[c] = floor(1.25 (c + 64)).
"""
from __future__ import annotations

import functools
import math

import torch


@functools.lru_cache(maxsize=64)
def resample_maps_host(H: int, W: int, crop_size: int = 384, strategy: str = "center_crop"):
    """(row_map, col_map) as tuples of ints, each crop_size long."""
    if strategy == "resize":
        new_h = new_w = crop_size
        top = left = 0
    elif strategy == "center_crop":
        new_h = crop_size
        new_w = int(W * (crop_size / H))
        left = (new_w - crop_size) // 2
        top = (new_h - crop_size) // 2
    else:
        raise ValueError(f"unknown strategy {strategy!r} (llava/video_utils.py:293-308 knows 'resize' and 'center_crop')")
    inv_y, inv_x = 1.0 / (new_h / H), 1.0 / (new_w / W)
    rows = tuple(min(int(math.floor(r * inv_y)), H - 1) for r in range(top, top + crop_size))
    cols = tuple(min(int(math.floor(c * inv_x)), W - 1) for c in range(left, left + crop_size))
    return rows, cols


@functools.lru_cache(maxsize=64)
def _maps_device(H, W, crop_size, strategy, device_str):
    rows, cols = resample_maps_host(H, W, crop_size, strategy)
    dev = torch.device(device_str)
    return (torch.tensor(rows, dtype=torch.int32, device=dev), torch.tensor(cols, dtype=torch.int32, device=dev))


def resample_maps(H: int, W: int, crop_size: int, strategy: str, device):
    """Device int32 tensors of the maps (cached per geometry and device)."""
    return _maps_device(H, W, crop_size, strategy, str(torch.device(device)))

"""PCIe host->device rates for the depth upload: strided window copy (v3d_upload_depth_window) vs one contiguous copy
of the same bytes vs the full frames.  python tools/time_h2d.py [frames]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__  # noqa: F401
from video_3d_llm_b200.frontend import FrontEnd

V = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda:0")
fe = FrontEnd(device=dev, patch_size=14, embedding_size=1152)
h = torch.zeros((V, 480, 640), dtype=torch.int16).pin_memory()
d = torch.empty((V, 480, 640), dtype=torch.int16, device=dev)
nb = fe.upload_bytes(V, 2)
hc = torch.zeros(nb, dtype=torch.uint8).pin_memory()
dc = torch.empty(nb, dtype=torch.uint8, device=dev)


def timed(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


t_win = timed(lambda: fe.upload(h, out=d))
t_con = timed(lambda: dc.copy_(hc, non_blocking=True))
t_full = timed(lambda: d.copy_(h, non_blocking=True))
print({"frames": V, "window_bytes": nb, "strided_window_ms": t_win, "strided_window_GBs": nb / t_win / 1e6,
       "contiguous_same_bytes_ms": t_con, "contiguous_GBs": nb / t_con / 1e6, "full_frames_ms": t_full,
       "full_GBs": h.numel() * 2 / t_full / 1e6})

"""Time the projector-tail kernel (k_pool_pe_grid) on the bench's workload: python tools/time_tail.py [frames] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
print(bench.tail_leg(torch.device("cuda:0"), frames, steps))

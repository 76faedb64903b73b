"""BASELINE configs[4] alone (bench.py --config5 runs it inside the bench): python tools/run_config5.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
print(json.dumps(bench.config5_leg(0, 1, dev)))

import sys, json, torch
sys.path.insert(0, '/root/repo')
import bench
dev = torch.device('cuda:0'); torch.cuda.set_device(dev)
print(json.dumps(bench.tail_commute_leg(dev)))

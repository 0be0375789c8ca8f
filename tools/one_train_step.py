"""One vgg-Q training step (for ncu): warm-up + 1 profiled `DreamNetwork.train` at batch B."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dream_b200 import network
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
net = network.create_network_from_config_data(bench.make_config()); net.enable_training()
x = torch.rand((B, 3, 400, 400), device="cuda") * 2 - 1
t = torch.rand((B, 7, 100, 100), device="cuda")
for _ in range(2):
    net.train([x], t)
torch.cuda.synchronize()

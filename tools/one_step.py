"""One bench step (for ncu): `python tools/one_step.py [batch]` runs warm-up + 1 forward+peaks step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dream_b200 import network, image_proc
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
net = network.create_network_from_config_data(bench.make_config()); net.enable_evaluation()
x = torch.rand((B, 3, 400, 400), device="cuda") * 2 - 1
for _ in range(2):
    with torch.no_grad():
        t = image_proc.find_peaks_device(net.model.module.belief_maps(x), 0.4395)
        k = image_proc.select_keypoints_device(t, 0.25)
torch.cuda.synchronize()

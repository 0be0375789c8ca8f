"""Per-parameter gradient error of the hand-written backward vs the oracle's autograd (diagnostic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from oracle import ref_models
from dream_b200 import models

mode = sys.argv[1] if len(sys.argv) > 1 else "default"
g = np.load(os.path.join(ROOT, "tests", "golden", "net_vgg_q%s.npz" % ("" if mode == "default" else "_he")))
shapes = ref_models.vgg_state_shapes(7, prefix="")
sd = ref_models.synth_state_dict(shapes, seed=0, out_gain=float(g["gain"]), mode=mode)
x = torch.from_numpy(g["x"]); target = torch.from_numpy(g["target"])
osd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
loss = torch.nn.functional.mse_loss(ref_models.vgg_forward(osd, x, prefix=""), target)
loss.backward()
net = models.DreamHourglass(7, internalize_spatial_softmax=False); net.load_state_dict(sd); net = net.cuda().train()
out = net(x.cuda())[0]
l2 = torch.nn.MSELoss()(out, target.cuda()); l2.backward()
print("loss ref %.6g ours %.6g" % (loss.item(), l2.item()))
for name, p in net.named_parameters():
    ref = osd[name].grad
    got = p.grad.cpu()
    rel = float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
    cos = float((got * ref).sum() / (got.norm() * ref.norm()).clamp_min(1e-30))
    print("%-28s rel %.4f cos %.6f |ref| %.3g |got| %.3g" % (name, rel, cos, ref.norm().item(), got.norm().item()))


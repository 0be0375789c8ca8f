"""Per-parameter gradient diagnostics for ResnetSimple training (diagnostic; prints cos / rel per parameter)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import ref_models
from dream_b200 import models
full = len(sys.argv) > 1 and sys.argv[1] == "full"
shapes = ref_models.resnet_state_shapes(7, full=full, prefix="")
sd = ref_models.synth_state_dict(shapes, seed=3, out_gain=0.04, mode="he")
gen = torch.Generator().manual_seed(5)
x = torch.rand((2, 3, 96, 80), generator=gen) * 2 - 1
osd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
y = ref_models.resnet_forward(osd, x, full=full, training=True, prefix="")
target = torch.rand(y.shape, generator=gen)
rl = torch.nn.functional.mse_loss(y, target); rl.backward()
net = models.ResnetSimple(7, full=full); net.load_state_dict(sd); net = net.cuda().train()
out = net(x.cuda())[0]
print("fwd max abs diff", (out.cpu() - y.detach()).abs().max().item(), "ref max", y.abs().max().item())
l = torch.nn.MSELoss()(out, target.cuda()); l.backward()
print("loss ref %.6g ours %.6g" % (rl.item(), l.item()))
rows = []
for name, p in net.named_parameters():
    ref = osd[name].grad; got = p.grad.cpu()
    cos = float((got * ref).sum() / (got.norm() * ref.norm()).clamp_min(1e-30))
    rows.append((name, cos, float(got.norm() / ref.norm().clamp_min(1e-30))))
for r in rows[:6] + rows[-14:]:
    print("%-34s cos %.5f norm ratio %.4f" % r)
worst = sorted(rows, key=lambda r: r[1])[:8]
print("worst:"); [print("  %-34s cos %.5f norm ratio %.4f" % r) for r in worst]

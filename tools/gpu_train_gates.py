"""Measures the quantities the training tests gate on (run on a GPU box; prints one JSON line per experiment):
  resnet  : per-parameter gradient cosine / norm ratio vs the oracle's fp32 autograd for several batch shapes
  determ  : two identical training forward/backward passes -> are loss / BN statistics / gradients bit-identical?
  curve   : N optimizer steps through DreamNetwork.train vs the same steps on the oracle (vgg-Q)
  huber   : DreamNetwork.loss with the SmoothL1 criterion vs the oracle
python tools/gpu_train_gates.py [resnet|determ|curve|huber ...]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import ref_models


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(a @ b / (a.norm() * b.norm()).clamp_min(1e-300))


def resnet_case(full, shape, seed=3):
    from dream_b200 import models
    shapes = ref_models.resnet_state_shapes(7, full=full, prefix="")
    sd = ref_models.synth_state_dict(shapes, seed=seed, out_gain=0.04, mode="he")
    gen = torch.Generator().manual_seed(5)
    x = torch.rand(shape, generator=gen) * 2 - 1
    net = models.ResnetSimple(7, full=full)
    net.load_state_dict(sd)
    net = net.cuda().train()
    osd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
           for k, v in sd.items()}
    t0 = time.time()
    y = ref_models.resnet_forward(osd, x, full=full, training=True, prefix="")
    target = torch.rand(y.shape, generator=gen)
    ref_loss = torch.nn.functional.mse_loss(y, target)
    ref_loss.backward()
    t_oracle = time.time() - t0
    out = net(x.cuda())[0]
    loss = torch.nn.MSELoss()(out, target.cuda())
    loss.backward()
    rows = []
    for name, p in net.named_parameters():
        ref = osd[name].grad
        if ref.norm() < 1e-12 * max(1.0, float(p.grad.norm())) or (name.startswith("upsample") and name.endswith("bias") and "12" not in name and "2.3" not in name):
            continue
        rows.append((name, cos(p.grad.cpu(), ref), float(p.grad.cpu().norm() / ref.norm())))
    trunk = [r for r in rows if not r[0].startswith("upsample")]
    dec = [r for r in rows if r[0].startswith("upsample")]
    rat = sorted(r[2] for r in rows)
    return {"exp": "resnet", "full": full, "shape": list(shape), "oracle_s": round(t_oracle, 1),
            "loss_rel": abs(loss.item() - ref_loss.item()) / ref_loss.item(),
            "trunk_min_cos": min(r[1] for r in trunk), "trunk_worst": min(trunk, key=lambda r: r[1])[0],
            "trunk_ratio_range": [min(r[2] for r in trunk), max(r[2] for r in trunk)],
            "dec_min_cos": min(r[1] for r in dec), "dec_ratio_range": [min(r[2] for r in dec), max(r[2] for r in dec)],
            "median_ratio": rat[len(rat) // 2],
            "cos_below_0.985": sum(1 for r in rows if r[1] < 0.985), "n": len(rows)}


def determ_case():
    from dream_b200 import models
    shapes = ref_models.resnet_state_shapes(7, full=False, prefix="")
    sd = ref_models.synth_state_dict(shapes, seed=3, out_gain=0.04, mode="he")
    gen = torch.Generator().manual_seed(5)
    x = (torch.rand((2, 3, 160, 160), generator=gen) * 2 - 1).cuda()
    outs = []
    for _ in range(2):
        net = models.ResnetSimple(7, full=False)
        net.load_state_dict(sd)
        net = net.cuda().train()
        out = net(x)[0]
        loss = out.pow(2).mean()
        loss.backward()
        outs.append((out.detach().clone(), {k: v.clone() for k, v in net.named_buffers()},
                     {k: p.grad.clone() for k, p in net.named_parameters()}))
    a, b = outs
    same_g = [k for k in a[2] if torch.equal(a[2][k], b[2][k])]
    bn_g = [k for k in a[2] if ("bn" in k or k.split(".")[-2].isdigit() and a[2][k].dim() == 1)]
    return {"exp": "determ", "output_identical": bool(torch.equal(a[0], b[0])),
            "buffers_identical": all(torch.equal(a[1][k], b[1][k]) for k in a[1]),
            "grads_identical": len(same_g), "grads_total": len(a[2]),
            "bn_grads_identical": sum(1 for k in bn_g if k in same_g), "bn_grads_total": len(bn_g),
            "max_grad_rel_diff": max(float((a[2][k] - b[2][k]).abs().max() / a[2][k].abs().max().clamp_min(1e-30)) for k in a[2])}


def curve_case(res, B, steps, lr, opt_type="sgd", loss_type="mse", mode="he", gain=0.1):
    from conftest import panda_config
    from dream_b200 import network
    cfg = panda_config("vgg")
    cfg["architecture"]["loss"] = {"type": loss_type}
    cfg["training"]["config"]["net_input_resolution"] = [res[1], res[0]]
    cfg["training"]["config"]["optimizer"] = {"type": opt_type, "learning_rate": lr}
    net = network.create_network_from_config_data(cfg)
    sd = ref_models.synth_state_dict(ref_models.vgg_state_shapes(7), seed=5, out_gain=gain, mode=mode)
    net.model.load_state_dict(sd)
    net.enable_training()
    gen = torch.Generator().manual_seed(1)
    x = torch.rand((B, 3, res[0], res[1]), generator=gen) * 2 - 1
    out_w, out_h = net.trained_net_output_resolution()
    t = torch.rand((B, 7, out_h, out_w), generator=gen)
    osd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt = (torch.optim.SGD if opt_type == "sgd" else torch.optim.Adam)(list(osd.values()), lr=lr)
    crit = torch.nn.MSELoss() if loss_type == "mse" else torch.nn.SmoothL1Loss()
    ours, refs = [], []
    t0 = time.time()
    for step in range(steps):
        ours.append(net.train([x.cuda()], t.cuda()).item())
        opt.zero_grad()
        ref = crit(ref_models.vgg_forward(osd, x), t)
        ref.backward()
        opt.step()
        refs.append(ref.item())
    rel = [abs(a - b) / b for a, b in zip(ours, refs)]
    return {"exp": "curve", "res": list(res), "B": B, "steps": steps, "opt": opt_type, "lr": lr, "loss": loss_type,
            "secs": round(time.time() - t0, 1), "first": refs[0], "last": refs[-1], "last_ours": ours[-1],
            "max_rel": max(rel), "rel_last": rel[-1]}


def huber_case():
    from conftest import panda_config
    from dream_b200 import network
    cfg = panda_config("vgg")
    cfg["architecture"]["loss"] = {"type": "huber"}
    cfg["training"]["config"]["net_input_resolution"] = [96, 64]
    net = network.create_network_from_config_data(cfg)
    sd = ref_models.synth_state_dict(ref_models.vgg_state_shapes(7), seed=6, out_gain=13.0, mode="default")
    net.model.load_state_dict(sd)
    net.enable_training()
    gen = torch.Generator().manual_seed(2)
    x = torch.rand((3, 3, 64, 96), generator=gen) * 2 - 1
    t = torch.rand((3, 7, 16, 24), generator=gen) * 3 - 1          # |error| on both sides of the Huber knee
    net.optimizer.zero_grad()
    loss = net.loss([x.cuda()], t.cuda())
    loss.backward()
    osd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = torch.nn.SmoothL1Loss()(ref_models.vgg_forward(osd, x), t)
    ref.backward()
    rows = [(n, cos(p.grad.cpu(), osd[n].grad), float(p.grad.cpu().norm() / osd[n].grad.norm()))
            for n, p in net.model.named_parameters() if n.endswith("weight")]
    hk = "module.heads_0.4.weight"
    g = dict(net.model.named_parameters())[hk].grad.cpu()
    return {"exp": "huber", "loss_rel": abs(loss.item() - ref.item()) / ref.item(),
            "min_cos": min(r[1] for r in rows), "ratio_range": [min(r[2] for r in rows), max(r[2] for r in rows)],
            "head_rel": float((g - osd[hk].grad).abs().max() / osd[hk].grad.abs().max())}


if __name__ == "__main__":
    torch.backends.cudnn.allow_tf32 = False
    which = sys.argv[1:] or ["resnet", "determ", "curve", "huber"]
    if "resnet" in which:
        for full, shape in ((False, (2, 3, 160, 160)), (False, (4, 3, 224, 224)), (False, (8, 3, 192, 192)),
                            (True, (2, 3, 160, 160)), (True, (4, 3, 160, 192))):
            print(json.dumps(resnet_case(full, shape)), flush=True)
    if "determ" in which:
        print(json.dumps(determ_case()), flush=True)
    if "curve" in which:
        print(json.dumps(curve_case((64, 96), 4, 4, 0.002)), flush=True)
        print(json.dumps(curve_case((192, 192), 2, 20, 0.002)), flush=True)
        print(json.dumps(curve_case((192, 192), 2, 20, 1.5e-4, "adam", mode="default", gain=13.0)), flush=True)
        print(json.dumps(curve_case((400, 400), 2, 12, 0.002)), flush=True)
    if "huber" in which:
        print(json.dumps(huber_case()), flush=True)

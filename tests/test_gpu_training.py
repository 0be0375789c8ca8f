"""GPU: training path (forward with saved activations + hand-written backward) against the oracle's
autograd and the reference golden gradients.  Gate: per-parameter gradient error
<= 1e-2 relative (fp16 operands), loss within 1e-3 relative."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ref_models  # noqa: E402


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def _oracle_grads(sd, x, target, **kw):
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    y = ref_models.vgg_forward(sd, x, prefix="", **kw)
    loss = torch.nn.functional.mse_loss(y, target)
    loss.backward()
    return loss.item(), {k: v.grad for k, v in sd.items()}, y.detach()


def _wgrad_reference(dy, x, taps):
    import torch.nn.functional as F
    B, H, W, Ci = x.shape
    xp = F.pad(x.float(), (0, 0, 1, 1, 1, 1))
    ref = torch.empty((len(taps), dy.shape[3], Ci), device=x.device)
    for t, (dyy, dxx) in enumerate(taps):
        ref[t] = torch.einsum("bhwo,bhwi->oi", dy.float(), xp[:, 1 + dyy:1 + dyy + H, 1 + dxx:1 + dxx + W, :])
    return ref


def test_wgrad_kernel_matches_torch(built_lib):
    from dream_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    for (B, H, W, Ci, Co) in [(2, 16, 24, 64, 64), (3, 25, 25, 128, 256), (2, 50, 37, 256, 128)]:
        x = (torch.randn((B, H, W, Ci), device="cuda", generator=g) * 0.5).half()
        dy = (torch.randn((B, H, W, Co), device="cuda", generator=g) * 0.5).half()
        dw = ops.wgrad(dy, x, ops.TAPS_3x3)                         # [9, Co, Ci]
        ref = _wgrad_reference(dy, x, ops.TAPS_3x3)
        assert _rel(dw, ref) <= 2e-3, (B, H, W, Ci, Co, _rel(dw, ref))


def test_streaming_backward_kernels_match_torch(built_lib):
    import torch.nn.functional as F
    from dream_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randn((2, 26, 31, 64), device="cuda", generator=g).half()
    xr = x.permute(0, 3, 1, 2).float().requires_grad_(True)
    # max-pool backward (floor mode, odd width)
    yp = F.max_pool2d(xr, 2)
    dy = torch.randn(yp.shape, device="cuda", generator=g).half()
    yp.backward(dy.float())
    dx = ops.maxpool2_bwd(x, dy.permute(0, 2, 3, 1).contiguous())
    assert torch.equal(dx.permute(0, 3, 1, 2).float(), xr.grad)
    # upsample backward
    xr.grad = None
    yu = F.interpolate(xr, scale_factor=2)
    dyu = torch.randn(yu.shape, device="cuda", generator=g).half()
    yu.backward(dyu.float())
    dxu = ops.upsample2_bwd(dyu.permute(0, 2, 3, 1).contiguous())
    assert (dxu.permute(0, 3, 1, 2).float() - xr.grad).abs().max() <= 4e-3 * xr.grad.abs().max()
    # relu mask + bias grad
    y = torch.relu(x)
    d = torch.randn(x.shape, device="cuda", generator=g).half()
    ref_mask = d.float() * (y.float() > 0)
    got = ops.relu_mask_(d.clone(), y)
    assert torch.equal(got.float(), ref_mask)
    db = ops.bias_grad(got)
    assert (db - ref_mask.sum(dim=(0, 1, 2))).abs().max() <= 1e-3 * ref_mask.abs().sum(dim=(0, 1, 2)).max()


@pytest.mark.parametrize("fixture", ["vgg_q", "vgg_q_he"])
def test_hourglass_gradients_match_oracle_and_golden(fixture, golden_dir, built_lib):
    from dream_b200 import models
    g = np.load(os.path.join(golden_dir, "net_%s.npz" % fixture))
    shapes = ref_models.vgg_state_shapes(7, prefix="")
    sd = ref_models.synth_state_dict(shapes, seed=0, out_gain=float(g["gain"]), mode=str(g["mode"]))
    x = torch.from_numpy(g["x"])
    target = torch.from_numpy(g["target"])
    ref_loss, ref_grads, _ = _oracle_grads(sd, x, target)
    assert abs(ref_loss - float(g["loss"])) <= 1e-5 * float(g["loss"])

    net = models.DreamHourglass(7, internalize_spatial_softmax=False)
    net.load_state_dict(sd)
    net = net.cuda().train()
    out = net(x.cuda())[0]
    loss = torch.nn.MSELoss()(out, target.cuda())
    loss.backward()
    assert abs(loss.item() - ref_loss) <= 1e-3 * ref_loss
    worst = 0.0
    for name, p in net.named_parameters():
        assert p.grad is not None, name
        r = _rel(p.grad.cpu(), ref_grads[name])
        worst = max(worst, r)
        assert r <= 1e-2, (name, r)
    print("worst per-parameter grad rel err:", worst)
    for key in g.files:
        if key.startswith("grad::"):
            ref = torch.from_numpy(g[key])
            got = dict(net.named_parameters())[key[6:]].grad.cpu()[:ref.shape[0]]
            assert _rel(got, ref) <= 1e-2, key


def test_training_steps_track_oracle_loss_curve(built_lib):
    """A few SGD steps through DreamNetwork.train vs the same steps on the oracle (same seed/weights)."""
    from conftest import panda_config
    from dream_b200 import network
    cfg = panda_config("vgg")
    cfg["training"]["config"]["net_input_resolution"] = [96, 64]
    cfg["training"]["config"]["optimizer"] = {"type": "sgd", "learning_rate": 0.002}
    net = network.create_network_from_config_data(cfg)
    sd = ref_models.synth_state_dict(ref_models.vgg_state_shapes(7), seed=5, out_gain=0.1, mode="he")
    net.model.load_state_dict(sd)
    net.enable_training()
    gen = torch.Generator().manual_seed(1)
    x = torch.rand((4, 3, 64, 96), generator=gen) * 2 - 1
    t = torch.rand((4, 7, 16, 24), generator=gen)
    osd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.SGD(list(osd.values()), lr=0.002)
    for step in range(4):
        loss = net.train([x.cuda()], t.cuda())
        opt.zero_grad()
        ref = torch.nn.functional.mse_loss(ref_models.vgg_forward(osd, x), t)
        ref.backward()
        opt.step()
        assert abs(loss.item() - ref.item()) <= 5e-3 * ref.item(), (step, loss.item(), ref.item())

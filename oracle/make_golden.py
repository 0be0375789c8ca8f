"""Generate tests/golden/*.npz by running the REFERENCE's own code (build container only).

/root/reference cannot travel to the GPU box, so its outputs are pinned here as small fixtures:
  * the reference nn.Modules (dream/models.py, loaded by file path under a stub `dream` package;
    torchvision constructors patched to weights=None because pretrained=True would download,
    models.py:22,587) evaluated on CPU fp32 with oracle.ref_models.synth_state_dict weights
    (regenerable anywhere from key names, so the weights are NOT stored);
  * the reference peaks_from_belief_maps (dream/image_proc.py:914-1018, imported with matplotlib /
    webcolors stubbed) on synthetic maps, incl. the reference's own test_belief_maps case;
  * the reference SoftArgmaxPavlo (dream/spatial_softmax.py).
Run:  python oracle/make_golden.py        (requires /root/reference)
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("DREAM_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")

from oracle import ref_models, ref_peaks  # noqa: E402


def load_reference():
    """Import dream.models / dream.spatial_softmax / dream.image_proc from the reference tree."""
    import torchvision.models as tvm
    pkg = types.ModuleType("dream")
    pkg.__path__ = [os.path.join(REF, "dream")]
    sys.modules["dream"] = pkg
    for stub in ("matplotlib", "matplotlib.pyplot", "webcolors"):
        sys.modules.setdefault(stub, types.ModuleType(stub))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

    def load(name):
        spec = importlib.util.spec_from_file_location("dream." + name, os.path.join(REF, "dream", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["dream." + name] = mod
        spec.loader.exec_module(mod)
        setattr(pkg, name, mod)
        return mod
    # SoftArgmaxPavlo calls .cuda() for its index grids (spatial_softmax.py:23,66,79): make it a no-op on CPU
    torch.Tensor.cuda = lambda self, *a, **k: self
    orig_vgg, orig_res = tvm.vgg19, tvm.resnet101
    tvm.vgg19 = lambda pretrained=False, **k: orig_vgg(weights=None)
    tvm.resnet101 = lambda pretrained=False, **k: orig_res(weights=None)
    return load("spatial_softmax"), load("models"), load("image_proc")


def main():
    os.makedirs(GOLD, exist_ok=True)
    softmax_mod, models, image_proc = load_reference()
    torch.set_num_threads(8)

    # ---------------- networks ----------------
    # name -> (constructor, shape table, input shape, weight mode).  "default" = the reference's own
    # initialisation statistics (the 1e-3 gate); "he" = stress weights (see ref_models).
    V = ref_models.vgg_state_shapes
    R = ref_models.resnet_state_shapes
    HG = models.DreamHourglass
    nets = {
        "vgg_q": (lambda: HG(7, internalize_spatial_softmax=False), V(7, prefix=""), (2, 3, 64, 80), "default"),
        "vgg_q_he": (lambda: HG(7, internalize_spatial_softmax=False), V(7, prefix=""), (2, 3, 64, 80), "he"),
        "vgg_f": (lambda: HG(7, internalize_spatial_softmax=False, deconv_decoder=True, full_output=True),
                  V(7, deconv_decoder=True, full_output=True, prefix=""), (2, 3, 48, 64), "default"),
        "vgg_f_he": (lambda: HG(7, internalize_spatial_softmax=False, deconv_decoder=True, full_output=True),
                     V(7, deconv_decoder=True, full_output=True, prefix=""), (1, 3, 48, 64), "he"),
        "vgg_q_skip": (lambda: HG(7, internalize_spatial_softmax=False, skip_connections=True),
                       V(7, prefix=""), (1, 3, 64, 64), "default"),
        "vgg_q_full": (lambda: HG(7, internalize_spatial_softmax=False, full_output=True),
                       V(7, full_output=True, prefix=""), (1, 3, 48, 48), "default"),
        "resnet_h_he": (lambda: models.ResnetSimple(7, pretrained=False), R(7, prefix=""), (2, 3, 96, 80), "he"),
        "resnet_f_he": (lambda: models.ResnetSimple(7, pretrained=False, full=True), R(7, full=True, prefix=""),
                        (1, 3, 72, 104), "he"),
    }
    for name, (ctor, shapes, xshape, mode) in nets.items():
        net = ctor().eval()
        ref_sd = net.state_dict()
        assert {k: tuple(v.shape) for k, v in ref_sd.items()} == {k: tuple(s) for k, s in shapes.items()}, \
            "oracle shape table disagrees with the reference module for " + name
        assert list(ref_sd.keys()) == list(shapes.keys()), "key order differs for " + name
        g = torch.Generator().manual_seed(42)
        x = torch.rand(xshape, generator=g) * 2 - 1
        # choose the output gain so the belief maps peak at ~1 like a trained network's
        net.load_state_dict(ref_models.synth_state_dict(shapes, seed=0, out_gain=1.0, mode=mode))
        with torch.no_grad():
            gain = float(np.float32(1.0 / net(x)[0].abs().max().item()))
        sd = ref_models.synth_state_dict(shapes, seed=0, out_gain=gain, mode=mode)
        net.load_state_dict(sd)
        with torch.no_grad():
            y = net(x)[0]
        out = {"x": x.numpy(), "y": y.numpy(), "gain": np.float64(gain), "mode": np.array(mode)}
        if name in ("vgg_q", "vgg_q_he", "resnet_h_he"):
            # gradients of an MSE loss (network.py:350-359) w.r.t. a few parameters
            net.train()                      # BN uses batch statistics in training (resnet)
            net.zero_grad()
            tg = torch.rand(y.shape, generator=g)
            yt = net(x)[0]
            loss = torch.nn.MSELoss()(yt, tg)
            loss.backward()
            out["target"] = tg.numpy()
            out["loss"] = np.float64(loss.item())
            params = dict(net.named_parameters())
            picks = [k for k in params if k.endswith("weight")]
            picks = [picks[0], picks[1], picks[len(picks) // 2], picks[-2], picks[-1]]
            for k in picks:
                gk = params[k].grad.numpy()
                out["grad::" + k] = gk if gk.size <= 100000 else gk[:4]     # leading slice keeps fixtures small
        np.savez_compressed(os.path.join(GOLD, "net_%s.npz" % name), **out)
        print(name, mode, tuple(y.shape), "gain %.4g" % gain, float(y.abs().max()), float(y.std()))

    # ---------------- peaks ----------------
    rng = np.random.default_rng(7)
    maps = []
    # (0,1) the reference's own known-answer test: test/test_image_proc.py:94-120
    maps.append(("ref_test", image_proc.create_belief_map((80, 60), [np.array([65.0, 20.0]),
                                                                       np.array([100.0, 80.0])]).astype(np.float32)))
    # single gaussians at random (incl. near-border) positions, 100x100
    pts = [(rng.uniform(0, 100), rng.uniform(0, 100)) for _ in range(12)] + [(2.5, 2.2), (97.9, 50.0), (50.0, 4.0)]
    maps.append(("single100", image_proc.create_belief_map((100, 100), pts).astype(np.float32)))
    # two-peak maps with score gaps around the 0.25 decision threshold
    two = []
    for gap in (0.0, 0.2, 0.2499, 0.25, 0.2501, 0.3, 0.6):
        a = image_proc.create_belief_map((100, 100), [(30.3, 40.7)])[0]
        b = image_proc.create_belief_map((100, 100), [(70.1, 60.2)])[0]
        two.append((a + (1.0 - gap) * b).astype(np.float32))
    maps.append(("two100", np.stack(two)))
    # noise, plateaus, negatives, constant maps
    noisy = [rng.standard_normal((100, 100)).astype(np.float32) * s for s in (0.05, 0.3, 1.0)]
    plate = np.zeros((100, 100), np.float32); plate[40:60, 30:70] = 0.8
    neg = -np.abs(rng.standard_normal((100, 100))).astype(np.float32)
    const = np.full((100, 100), 0.5, np.float32)
    ties = np.zeros((100, 100), np.float32); ties[20, 20] = 1.0; ties[20, 60] = 1.0; ties[70, 40] = 1.0
    maps.append(("misc100", np.stack(noisy + [plate, neg, const, ties])))
    # non-square, small and large maps
    maps.append(("rect", (rng.random((3, 52, 77)).astype(np.float32) ** 6)))
    maps.append(("tiny", (rng.random((4, 9, 7)).astype(np.float32))))
    big = image_proc.create_belief_map((400, 400), [(rng.uniform(5, 395), rng.uniform(5, 395)) for _ in range(3)])
    maps.append(("big400", (big + 0.02 * rng.standard_normal(big.shape)).astype(np.float32)))
    out = {}
    for name, m in maps:
        for off in (0.0, 0.4395):
            pk = image_proc.peaks_from_belief_maps(torch.from_numpy(m), off)
            flat = [(j, p[0], p[1], float(p[2]), p[3]) for j, lst in enumerate(pk) for p in lst]
            out["%s::maps" % name] = m
            out["%s::peaks@%g" % (name, off)] = np.array(flat, dtype=np.float64).reshape(-1, 5)
        print("peaks", name, m.shape, len(flat))
    np.savez_compressed(os.path.join(GOLD, "peaks.npz"), **out)

    # ---------------- soft-argmax ----------------
    hm = torch.from_numpy(rng.standard_normal((2, 7, 50, 60)).astype(np.float32))
    sa = softmax_mod.SoftArgmaxPavlo(n_keypoints=7, learned_beta=True, initial_beta=25.0)
    with torch.no_grad():
        sa.beta.copy_(torch.linspace(1.0, 30.0, 7))
        xy = sa(hm)
    np.savez_compressed(os.path.join(GOLD, "softargmax.npz"), maps=hm.numpy(), beta=sa.beta.detach().numpy(),
                        xy=xy.numpy())
    print("softargmax", tuple(xy.shape))


if __name__ == "__main__":
    main()

"""CPU: pin the oracle restatement against golden vectors produced by the reference's own code
(oracle/make_golden.py) and against the reference's known-answer test for peak extraction."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_models, ref_peaks

NETS = {   # fixture -> (family, forward kwargs)
    "vgg_q": ("vgg", {}),
    "vgg_q_he": ("vgg", {}),
    "vgg_f": ("vgg", dict(deconv_decoder=True, full_output=True)),
    "vgg_f_he": ("vgg", dict(deconv_decoder=True, full_output=True)),
    "vgg_q_skip": ("vgg", dict(skip_connections=True)),
    "vgg_q_full": ("vgg", dict(full_output=True)),
    "resnet_h_he": ("resnet", dict(full=False)),
    "resnet_f_he": ("resnet", dict(full=True)),
}


def _state(name, g):
    kind, kw = NETS[name]
    if kind == "vgg":
        shapes = ref_models.vgg_state_shapes(7, deconv_decoder=kw.get("deconv_decoder", False),
                                             full_output=kw.get("full_output", False), prefix="")
    else:
        shapes = ref_models.resnet_state_shapes(7, full=kw["full"], prefix="")
    return ref_models.synth_state_dict(shapes, seed=0, out_gain=float(g["gain"]), mode=str(g["mode"]))


@pytest.mark.parametrize("name", sorted(NETS))
def test_model_restatement_matches_reference(name, golden_dir):
    g = np.load(os.path.join(golden_dir, "net_%s.npz" % name))
    sd = _state(name, g)
    x = torch.from_numpy(g["x"])
    kind, kw = NETS[name]
    with torch.no_grad():
        if kind == "vgg":
            y = ref_models.vgg_forward(sd, x, prefix="", **kw)
        else:
            y = ref_models.resnet_forward(sd, x, prefix="", **kw)
    assert y.shape == g["y"].shape
    # same ATen ops on the same machine class: allow only accumulation-order noise
    assert np.abs(y.numpy() - g["y"]).max() <= 2e-5 * max(1.0, np.abs(g["y"]).max())


@pytest.mark.parametrize("name", ["vgg_q", "vgg_q_he", "resnet_h_he"])
def test_model_gradients_match_reference(name, golden_dir):
    g = np.load(os.path.join(golden_dir, "net_%s.npz" % name))
    sd = _state(name, g)
    for v in sd.values():
        if v.is_floating_point() and not v.dim() == 0:
            v.requires_grad_(True)
    for k in list(sd):
        if k.endswith("running_mean") or k.endswith("running_var"):
            sd[k] = sd[k].detach().clone()
    x = torch.from_numpy(g["x"])
    kind, kw = NETS[name]
    if kind == "vgg":
        y = ref_models.vgg_forward(sd, x, prefix="", **kw)
    else:
        y = ref_models.resnet_forward(sd, x, prefix="", training=True, **kw)
    loss = torch.nn.functional.mse_loss(y, torch.from_numpy(g["target"]))
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * max(1.0, float(g["loss"]))
    loss.backward()
    for key in g.files:
        if key.startswith("grad::"):
            ref = g[key]
            got = sd[key[6:]].grad.numpy()[:ref.shape[0]]
            assert np.abs(got - ref).max() <= 1e-4 * max(1e-6, np.abs(ref).max()), key


def test_fp16_operand_emulation_is_close_to_fp32_and_restores_functional():
    import torch.nn.functional as F
    before = (F.conv2d, F.conv_transpose2d)
    sd = ref_models.synth_state_dict(ref_models.vgg_state_shapes(7, prefix=""), seed=0, out_gain=13.0,
                                     mode="default")
    x = torch.rand((1, 3, 32, 48), generator=torch.Generator().manual_seed(0)) * 2 - 1
    with torch.no_grad():
        ref = ref_models.vgg_forward(sd, x, prefix="")
        with ref_models.fp16_operands():
            emu = ref_models.vgg_forward(sd, x, prefix="")
    assert (F.conv2d, F.conv_transpose2d) == before
    err = (emu - ref).abs().max().item()
    assert 0 < err <= 1e-3 * max(1.0, ref.abs().max().item())


def test_shape_tables_param_counts():
    n = sum(int(np.prod(s)) for s in ref_models.vgg_state_shapes(7).values())
    assert n == 22220615                       # SURVEY.md 8a / BASELINE.md
    n = sum(int(np.prod(s)) for s in ref_models.vgg_state_shapes(7, True, True).values())
    assert n == 22442055
    sh = ref_models.resnet_state_shapes(7)
    n = sum(int(np.prod(s)) for k, s in sh.items() if not ("running" in k or "num_batches" in k))
    assert n == 54039367


def test_gaussian_restatement_is_bit_identical_to_scipy():
    import scipy.ndimage as ndi
    rng = np.random.default_rng(3)
    for shp in [(100, 100), (60, 80), (25, 25), (13, 17), (5, 7), (1, 30), (208, 208)]:
        m = rng.standard_normal(shp).astype(np.float32)
        assert np.array_equal(ref_peaks.gaussian_filter_f32(m), ndi.gaussian_filter(m, sigma=3))


def test_reference_known_answer_belief_maps():
    """The reference's own test (test/test_image_proc.py:94-120) run on the restatement."""
    res = (80, 60)
    kp = np.array([65.0, 20.0])
    maps = ref_peaks.create_belief_map(res, [kp, np.array([res[0] + 20.0, res[1] + 20.0])])
    peaks = ref_peaks.peaks_from_belief_maps(torch.tensor(maps).float().numpy(), 0.0)
    assert len(peaks[0]) == 1
    assert np.linalg.norm(kp - np.array(peaks[0][0][:2])) < 1.0e-3
    assert len(peaks[1]) == 0


def _golden_peak_sets(golden_dir):
    g = np.load(os.path.join(golden_dir, "peaks.npz"))
    names = sorted({k.split("::")[0] for k in g.files})
    return g, names


def test_peaks_restatement_matches_reference(golden_dir):
    g, names = _golden_peak_sets(golden_dir)
    for name in names:
        maps = g[name + "::maps"]
        for off in (0.0, 0.4395):
            ref = g["%s::peaks@%g" % (name, off)]
            got = ref_peaks.peaks_from_belief_maps(maps, off)
            flat = np.array([(j, p[0], p[1], float(p[2]), p[3]) for j, lst in enumerate(got) for p in lst],
                            dtype=np.float64).reshape(-1, 5)
            assert flat.shape == ref.shape, (name, off)
            assert np.array_equal(flat[:, [0, 3, 4]], ref[:, [0, 3, 4]]), (name, off)   # map id, score, peak id
            assert np.array_equal(flat[:, 1:3], ref[:, 1:3]), (name, off)               # refined x, y bit-exact


def test_create_belief_map_matches_reference_fixture(golden_dir):
    g, _ = _golden_peak_sets(golden_dir)
    maps = ref_peaks.create_belief_map((80, 60), [np.array([65.0, 20.0]), np.array([100.0, 80.0])])
    assert np.array_equal(maps.astype(np.float32), g["ref_test::maps"])


def test_select_keypoints_decision_table():
    pk = [[(1.0, 2.0, np.float32(0.9), 0)], [], [(1.0, 2.0, np.float32(0.9), 1), (3.0, 4.0, np.float32(0.6), 2)],
          [(1.0, 2.0, np.float32(0.9), 3), (3.0, 4.0, np.float32(0.7), 4)],
          [(5.0, 6.0, np.float32(0.5), 5), (7.0, 8.0, np.float32(0.75), 6)]]
    out = ref_peaks.select_keypoints(pk)
    assert out[0] == [1.0, 2.0]
    assert out[1] == [ref_peaks.SENTINEL] * 2
    assert out[2] == [1.0, 2.0]
    assert out[3] == [ref_peaks.SENTINEL] * 2
    assert out[4] == [7.0, 8.0]


def test_softargmax_restatement_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "softargmax.npz"))
    xy = ref_peaks.soft_argmax(g["maps"], g["beta"]).numpy()
    assert np.abs(xy - g["xy"]).max() <= 1e-4


# ------------------------------------------------------------------------------------------------
# host-side fixtures generated from the reference's image_proc / torchvision
# ------------------------------------------------------------------------------------------------
def test_input_oracle_matches_reference_fixtures(golden_dir):
    from oracle import ref_input
    GOLD = golden_dir
    g = np.load(os.path.join(GOLD, "normalize.npz"))
    for tag in ("half", "imagenet"):
        x = ref_input.normalize_u8(g["img"], g[tag + "::mean"], g[tag + "::std"])
        assert np.array_equal(x, g[tag + "::x"]), tag
    t = np.load(os.path.join(GOLD, "targets.npz"))
    ref = ref_input.create_belief_map((100, 100), t["pts"])
    assert np.array_equal(ref.astype(np.float32), t["tgt"]) and ref.sum() == float(t["tgt64_sum"])
    assert np.array_equal(ref_input.create_belief_map((208, 160), t["pts2"]).astype(np.float32), t["tgt2"])
    # the product's vectorised host version is the same function
    from dream_b200 import image_proc
    assert np.array_equal(image_proc.create_belief_map((100, 100), t["pts"]), ref)
    # window edge cases of the fixture: (4,4) and (94,94) stamp, (3.99,50), (95,50), (50.5,94.999), sentinel do not
    sums = t["tgt"].reshape(len(t["pts"]), -1).sum(1)
    assert sums[0] > 0 and sums[2] > 0 and sums[1] == 0 and sums[3] == 0 and sums[5] == 0


def test_frame_conversions_match_reference_fixture(golden_dir):
    GOLD = golden_dir
    from dream_b200 import analysis, image_proc
    from oracle import ref_analysis
    g = np.load(os.path.join(GOLD, "frames.npz"))
    n = len([k for k in g.files if k.endswith("::meta")])
    assert n == 5
    for i in range(n):
        preproc, rw, rh, nw, nh = g["case%d::meta" % i]
        raw, net_in = (int(rw), int(rh)), (int(nw), int(nh))
        net_out = (net_in[0] // 4, net_in[1] // 4)
        kps = g["case%d::kps" % i]
        netin = image_proc.convert_keypoints_to_netin_from_netout(kps, net_out, net_in)
        assert np.array_equal(netin, g["case%d::netin" % i])
        raw_kp = image_proc.convert_keypoints_to_raw_from_netin(netin, net_in, raw, preproc)
        assert np.array_equal(raw_kp, g["case%d::raw" % i])
        back = image_proc.convert_keypoints_to_netin_from_raw(raw_kp, raw, net_in, preproc)
        assert np.array_equal(back, g["case%d::back" % i])
        assert np.array_equal(analysis.detected_keypoints_raw(kps[None], net_out, net_in, raw, preproc)[0],
                              g["case%d::raw" % i])
        det, _ = ref_analysis.sample_loop(kps[None], np.zeros((1, len(kps), 2)), net_out, net_in, raw, preproc)
        assert np.array_equal(det[0], g["case%d::raw" % i])


@pytest.mark.parametrize("name,kw", [("ms2", dict(n_stages=2)), ("ms3_full", dict(n_stages=3, full_output=True))])
def test_multistage_restatement_matches_reference(name, kw, golden_dir):
    """DreamHourglassMultiStage (models.py:350-553): every stage's belief maps, the multi-stage loss of
    network.py:345-352 and gradients (incl. stage-1 parameters reached through later stages)."""
    g = np.load(os.path.join(golden_dir, "net_%s.npz" % name))
    S = kw["n_stages"]
    sd = ref_models.multistage_state_dict(7, S, g["gains"], full_output=kw.get("full_output", False), prefix="")
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    x = torch.from_numpy(g["x"])
    outs = ref_models.multistage_forward(sd, x, prefix="", **kw)
    assert len(outs) == S
    for s, y in enumerate(outs):
        assert np.abs(y.detach().numpy() - g["y%d" % (s + 1)]).max() <= 1e-5
    tg = torch.from_numpy(g["target"])
    loss = torch.nn.MSELoss()(torch.stack(outs), tg.unsqueeze(0).expand([S] + [-1] * tg.dim()))
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * float(g["loss"])
    loss.backward()
    for key in g.files:
        if key.startswith("grad::"):
            ref = g[key]
            got = sd[key[6:]].grad.numpy()[:ref.shape[0]]
            assert np.abs(got - ref).max() <= 1e-4 * max(np.abs(ref).max(), 1e-12), key

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

"""CPU: host-side logic of the drop-in (no GPU compute): C-ABI symbols, config handling, resolution
arithmetic (reference test/test_image_proc.py:20-91 expectations), tap tables, plan bookkeeping."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch


def test_capi_exports_every_declared_symbol(built_lib):
    from dream_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "dreamb200.h")).read()
    declared = set(re.findall(r"\b(dreamb200_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(built_lib)
    for sym in declared:
        assert hasattr(lib, sym), "libdreamb200.so does not export " + sym
    assert declared == set(_lib.DECLARED_SYMBOLS), declared ^ set(_lib.DECLARED_SYMBOLS)
    assert _lib.lib().dreamb200_version() >= 100


def test_conv_descriptor_layout_matches_header():
    from dream_b200._lib import ConvDesc
    # int8[16] x2 tap tables, pointer-aligned fields: catches drift between the ctypes mirror and the C struct
    assert ConvDesc.tap_dy.offset + 16 == ConvDesc.tap_dx.offset
    assert ctypes.sizeof(ConvDesc) % 8 == 0
    assert ConvDesc.y.offset % 8 == 0 and ConvDesc.residual.offset % 8 == 0


def test_shrink_and_crop_resolution_reference_expectations():
    from dream_b200 import image_proc
    # test/test_image_proc.py:20-91: (640,480) -> shrink (533,400); crop (480,480)@(80,0)
    assert image_proc.shrink_resolution((640, 480), (400, 400)) == (533, 400)
    assert image_proc.shrink_and_crop_resolution((640, 480), (400, 400)) == ((480, 480), (80, 0))
    assert image_proc.resolution_after_preprocessing((640, 480), (400, 400), "none") == (640, 480)
    assert image_proc.resolution_after_preprocessing((640, 480), (400, 400), "resize") == (400, 400)
    assert image_proc.resolution_after_preprocessing((640, 480), (400, 400), "shrink") == (533, 400)
    assert image_proc.resolution_after_preprocessing((640, 480), (400, 400), "shrink-and-crop") == (400, 400)
    with pytest.raises(AssertionError):
        image_proc.resolution_after_preprocessing((640, 480), (400, 400), "bogus")


def test_keypoint_frame_round_trips():
    from dream_b200 import image_proc
    kp = np.array([[10.0, 20.0], [55.5, 70.25]])
    for mode in image_proc.KNOWN_IMAGE_PREPROC_TYPES:
        netin_res = image_proc.resolution_after_preprocessing((640, 480), (400, 400), mode)
        a = image_proc.convert_keypoints_to_netin_from_raw(kp, (640, 480), netin_res, mode)
        b = image_proc.convert_keypoints_to_raw_from_netin(a, netin_res, (640, 480), mode)
        assert np.allclose(b, kp)
    out = image_proc.convert_keypoints_to_netin_from_netout(np.array([[50.0, 25.0]]), (100, 100), (400, 400))
    assert np.allclose(out, [[200.0, 100.0]])


def test_create_belief_map_matches_oracle():
    from dream_b200 import image_proc
    from oracle import ref_peaks
    pts = [(65.0, 20.0), (100.0, 80.0), (2.0, 2.0), (30.7, 40.2)]
    assert np.array_equal(image_proc.create_belief_map((80, 60), pts), ref_peaks.create_belief_map((80, 60), pts))


def test_gaussian_taps_match_oracle():
    from dream_b200 import image_proc
    from oracle import ref_peaks
    w, r = image_proc.gaussian_half_kernel()
    w2, r2 = ref_peaks.gaussian_weights()
    assert r == r2 == 12 and np.array_equal(w, w2)


def test_deconv_phase_taps_reproduce_conv_transpose():
    """The sub-pixel decomposition used for ConvTranspose2d (k3 s2 p1 op1 and k4 s2 p1) is exact."""
    from dream_b200.models import _deconv_phase_taps
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(0)
    for k, op in ((3, 1), (4, 0)):
        x = torch.randn((1, 5, 6, 7), generator=g)
        w = torch.randn((5, 4, k, k), generator=g)
        ref = F.conv_transpose2d(x, w, stride=2, padding=1, output_padding=op)
        out = torch.zeros_like(ref)
        H, W = x.shape[2:]
        assert ref.shape[2:] == (2 * H, 2 * W)
        xp = F.pad(x, (2, 2, 2, 2))
        for py in range(2):
            for px in range(2):
                acc = torch.zeros((1, 4, H, W))
                for dy, ky in _deconv_phase_taps(k, py):
                    for dx, kx in _deconv_phase_taps(k, px):
                        patch = xp[:, :, 2 + dy:2 + dy + H, 2 + dx:2 + dx + W]
                        acc += torch.einsum("bchw,co->bohw", patch, w[:, :, ky, kx])
                out[:, :, py::2, px::2] = acc
        assert torch.allclose(out, ref, atol=1e-4)


def test_parameter_trees_have_reference_keys():
    from dream_b200 import models
    from oracle import ref_models
    m = models.DataParallelShim(models.DreamHourglass(7, internalize_spatial_softmax=False))
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == \
        {k: tuple(s) for k, s in ref_models.vgg_state_shapes(7).items()}
    m = models.DataParallelShim(models.DreamHourglass(7, internalize_spatial_softmax=False, deconv_decoder=True,
                                                      full_output=True))
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == \
        {k: tuple(s) for k, s in ref_models.vgg_state_shapes(7, True, True).items()}
    for full in (False, True):
        m = models.DataParallelShim(models.ResnetSimple(7, full=full))
        assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == \
            {k: tuple(s) for k, s in ref_models.resnet_state_shapes(7, full=full).items()}
    n_params = sum(p.numel() for p in models.DreamHourglass(7, internalize_spatial_softmax=False).parameters())
    assert n_params == 22220615


def test_facade_validation_messages_without_gpu():
    from conftest import panda_config
    from dream_b200 import network
    cfg = panda_config()
    del cfg["architecture"]["type"]
    with pytest.raises(AssertionError, match='Required key "type"'):
        network.DreamNetwork(cfg)
    cfg = panda_config()
    cfg["architecture"]["type"] = "alexnet"
    with pytest.raises(AssertionError, match="known network architectures"):
        network.DreamNetwork(cfg)
    cfg = panda_config()
    cfg["training"]["config"]["net_input_resolution"] = [400]
    with pytest.raises(AssertionError, match="length 2"):
        network.DreamNetwork(cfg)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            network.DreamNetwork(panda_config())


def test_yaml_omap_loading(tmp_path):
    from dream_b200 import network
    p = tmp_path / "arch.yaml"
    p.write_text("!!omap\n- architecture: !!omap\n  - type: vgg\n  - input_heads:\n    - image_rgb\n"
                 "- training: !!omap\n  - config: !!omap\n    - net_input_resolution: [400, 400]\n")
    cfg = network.load_yaml_config(str(p))
    assert cfg["architecture"]["type"] == "vgg"
    assert cfg["training"]["config"]["net_input_resolution"] == [400, 400]
    out = tmp_path / "out.yaml"
    network.dump_yaml_config(cfg, str(out))
    assert network.load_yaml_config(str(out))["architecture"]["input_heads"] == ["image_rgb"]


# ------------------------------------------------------------------------------------------------
# batched post-loop of analyze_ndds_dataset vs the per-sample restatement
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("preproc,raw_res,net_in", [("shrink-and-crop", (640, 480), (400, 400)),
                                                    ("resize", (640, 480), (400, 400)),
                                                    ("none", (400, 400), (400, 400)),
                                                    ("shrink-and-crop", (480, 640), (400, 400))])
def test_batched_analysis_matches_sample_loop(preproc, raw_res, net_in):
    from dream_b200 import analysis
    from oracle import ref_analysis
    rng = np.random.default_rng(7)
    B, K = 33, 7
    net_out = (net_in[0] // 4, net_in[1] // 4)
    kps = rng.uniform(0, net_out[0], size=(B, K, 2)).astype(np.float32)
    kps[rng.random((B, K)) < 0.2] = -999.999                      # undetected
    kps[5] = -999.999                                              # a frame with no detection at all
    gt = rng.uniform(-60, max(raw_res) + 60, size=(B, K, 2))       # some out of frame
    gt[7, :, 0] = -5.0                                             # a frame with no in-frame ground truth
    gt[9, 0] = (0.0, 0.0); gt[9, 1] = (raw_res[0], raw_res[1])     # inclusive bounds
    det, metric = analysis.analyze_batch(kps, gt, net_out, net_in, raw_res, preproc)
    det_ref, metric_ref = ref_analysis.sample_loop(kps, gt, net_out, net_in, raw_res, preproc)
    assert det.shape == (B, K, 2) and det.dtype == np.float64
    assert np.array_equal(det, det_ref)                            # same float64 operations, same order
    assert np.array_equal(metric, metric_ref)
    assert metric[5] == 999.999 and metric[7] == 999.999


def test_pretrained_trunks_load_from_a_local_checkpoint_or_warn(tmp_path, monkeypatch):
    """dream/models.py:22,587 build the trunks from torchvision's pretrained nets; here that is a LOCAL file lookup
    (no download): found -> copied into the reference-named parameters, absent -> a warning, never silence."""
    import warnings
    from dream_b200 import models, pretrained
    monkeypatch.setenv("DREAMB200_PRETRAINED_DIR", str(tmp_path))
    monkeypatch.setattr(torch.hub, "get_dir", lambda: str(tmp_path / "nohub"))
    pretrained._WARNED.clear()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        models.DreamHourglass(7, internalize_spatial_softmax=False)
        models.ResnetSimple(7)
        models.ResnetSimple(7, pretrained=False)
    msgs = [str(x.message) for x in w]
    assert sum("vgg19" in m and "RANDOM" in m for m in msgs) == 1
    assert sum("resnet101" in m and "RANDOM" in m for m in msgs) == 1
    # a fake torchvision vgg19 checkpoint: every trunk conv except the fresh first one is taken from it
    net = models.DreamHourglass(7, internalize_spatial_softmax=False)
    fake = {}
    for name, p in net.named_parameters():
        if name.startswith("layer_0_"):
            idx, leaf = name.split(".")[1], name.split(".")[2]
            fake["features.%s.%s" % (idx, leaf)] = torch.full_like(p, float(idx) + (0.5 if leaf == "bias" else 0.0))
    torch.save(fake, tmp_path / "vgg19-dcbb9e9d.pth")
    net2 = models.DreamHourglass(7, internalize_spatial_softmax=False)
    sd = net2.state_dict()
    assert float(sd["layer_0_3_down.12.weight"].mean()) == 12.0 and float(sd["layer_0_5_down.34.bias"].mean()) == 34.5
    assert float(sd["layer_0_1_down.2.weight"].mean()) == 2.0
    assert float(sd["layer_0_1_down.0.weight"].abs().max()) < 1.0          # the fresh first conv keeps its own init


def test_invalidate_packed_weights_forgets_the_plan():
    """Updates that bypass torch's version counters (fused optimizers, raw pointer writes) need an explicit
    invalidation of the packed fp16 weights."""
    from dream_b200 import models
    net = models.DreamHourglass(7, internalize_spatial_softmax=False)
    net._plan, net._plan_key = {"stale": True}, ("key",)
    net.invalidate_packed_weights()
    assert net._plan is None and net._plan_key is None
    key0 = net._version_key()
    with torch.no_grad():
        next(net.parameters()).mul_(1.0)             # any in-place torch op bumps the counter the cache keys on
    assert net._version_key() != key0


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's CPU path = the oracle port, on the host cores) prints ONE JSON line
    with the CUDA arm's metric / unit / config plus impl, cpu_baseline and an e2e block without copies."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1"],
                         capture_output=True, text=True, timeout=900, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "images/sec" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert set(d["config"]) == {"workload", "parallelism", "l2"} and "batch 128/GPU, 400x400" in d["config"]["workload"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_two_row_accumulator_decomposition_is_the_convolution():
    """Executable statement of conv_rs3.cu's scheme: accumulator row (x, j) holds output rows 2j (first
    half of the columns) and 2j + 1 (second half); for input-row offset t = 0..3 and column tap s the A operand is the
    input pixel (2j - 1 + t, x - 1 + s) and the B operand is W[0,s] (t = 0, first half), [W[1,s] | W[0,s]] (t = 1),
    [W[2,s] | W[1,s]] (t = 2), W[2,s] (t = 3, second half).  Summed over (s, t) this is the 3x3 'same' convolution."""
    rng = np.random.default_rng(0)
    H, W, Ci, Co = 10, 7, 5, 3                                   # H even here; the kernel clips odd maps with TMA
    x = rng.standard_normal((H, W, Ci))
    w = rng.standard_normal((Co, Ci, 3, 3))
    xp = np.zeros((H + 2, W + 2, Ci)); xp[1:-1, 1:-1] = x         # zero padding = TMA out-of-bounds fill
    acc = np.zeros((H // 2, W, 2 * Co))                           # [j, x, (row 2j | row 2j + 1)]
    for s in range(3):
        for t in range(4):
            a = xp[t:t + H:2, s:s + W]                            # A[j, x] = input (2j - 1 + t, x - 1 + s), padded coords
            if t <= 2:
                acc[:, :, :Co] += a @ w[:, :, t, s].T             # output row 2j uses kernel row r = t
            if t >= 1:
                acc[:, :, Co:] += a @ w[:, :, t - 1, s].T         # output row 2j + 1 uses r = t - 1
    got = np.empty((H, W, Co)); got[0::2] = acc[:, :, :Co]; got[1::2] = acc[:, :, Co:]
    ref = np.zeros((H, W, Co))
    for r in range(3):
        for s in range(3):
            ref += xp[r:r + H, s:s + W] @ w[:, :, r, s].T
    assert np.allclose(got, ref, atol=1e-12)

"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path through the C-ABI against the
oracle restatement and the committed reference golden vectors.

Tolerances (BASELINE.json north_star): belief maps within 1e-3 max-abs of the CPU fp32 path at
trained-network scale (maps peaking near 1); integer peak coordinates bit-exact; refined (x, y)
within 1e-3 px (we in fact require exact equality with the reference fixture).
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ref_models, ref_peaks  # noqa: E402  (checker only)

NETS = {   # fixture -> (family, constructor / forward kwargs)
    "vgg_q": ("vgg", {}),
    "vgg_q_he": ("vgg", {}),
    "vgg_f": ("vgg", dict(deconv_decoder=True, full_output=True)),
    "vgg_f_he": ("vgg", dict(deconv_decoder=True, full_output=True)),
    "vgg_q_skip": ("vgg", dict(skip_connections=True)),
    "vgg_q_full": ("vgg", dict(full_output=True)),
    "resnet_h_he": ("resnet", dict(full=False)),
    "resnet_f_he": ("resnet", dict(full=True)),
}
# Gates.  (1) BELIEF_TOL = 1e-3 max-abs against the fp32 reference on belief maps scaled to peak at 1, with the
# reference's own initialisation statistics -- the north_star gate.  (2) "he" stress weights keep
# every layer's activations O(1) with no bias path to hide behind: there the 11-bit significand of ANY
# tensor-core operand format (fp16, and TF32 alike) costs 1.1e-3 .. 1.7e-3 by itself (CPU emulation
# oracle.ref_models.fp16_operands; two such emulations that differ only in accumulation order already sit that
# far apart after ~8 layers, because 1-ulp rounding flips cascade).  There the CUDA path must (a) stay within
# STRESS_TOL of fp32 and (b) be no further from fp32 than FLOOR_FACTOR x the emulation is, i.e. sit AT the
# operand format's noise floor.  Kernel exactness proper (one layer, same fp16 operands, error <= one fp16
# output ulp) is test_single_layers_are_exact_up_to_output_rounding.
BELIEF_TOL = 1e-3
STRESS_TOL = 2.5e-3
FLOOR_FACTOR = 1.5


def _shapes(name):
    kind, kw = NETS[name]
    if kind == "vgg":
        return ref_models.vgg_state_shapes(7, deconv_decoder=kw.get("deconv_decoder", False),
                                           full_output=kw.get("full_output", False), prefix="")
    return ref_models.resnet_state_shapes(7, full=kw["full"], prefix="")


def _build(name, sd):
    from dream_b200 import models
    kind, kw = NETS[name]
    if kind == "vgg":
        net = models.DreamHourglass(7, internalize_spatial_softmax=False, **kw)
    else:
        net = models.ResnetSimple(7, **kw)
    net.load_state_dict(sd, strict=True)
    return net.cuda().eval()


def _oracle(name, sd, x, emulate=False):
    kind, kw = NETS[name]
    with torch.no_grad():
        if emulate:
            with ref_models.fp16_operands():
                return _oracle(name, sd, x)
        if kind == "vgg":
            return ref_models.vgg_forward(sd, x, prefix="", **kw)
        return ref_models.resnet_forward(sd, x, prefix="", **kw)


@pytest.mark.parametrize("name", sorted(NETS))
def test_network_forward_matches_reference_golden(name, golden_dir, built_lib):
    g = np.load(os.path.join(golden_dir, "net_%s.npz" % name))
    mode = str(g["mode"])
    sd = ref_models.synth_state_dict(_shapes(name), seed=0, out_gain=float(g["gain"]), mode=mode)
    net = _build(name, sd)
    x = torch.from_numpy(g["x"])
    with torch.no_grad():
        y = net(x.cuda())[0].cpu().numpy()
    assert y.shape == g["y"].shape
    err = np.abs(y - g["y"]).max()
    emu = _oracle(name, sd, x, emulate=True).numpy()
    err_emu = np.abs(y - emu).max()
    print("%s [%s weights] max-abs vs reference fp32 golden %.3g, vs fp16-operand emulation %.3g (ref max %.3g)"
          % (name, mode, err, err_emu, np.abs(g["y"]).max()))
    floor = np.abs(emu - g["y"]).max()
    assert err <= (BELIEF_TOL if mode == "default" else STRESS_TOL)
    assert err <= FLOOR_FACTOR * floor + 1e-4, (err, floor)


@pytest.mark.parametrize("name,shape,mode", [
    ("vgg_q", (3, 3, 400, 400), "default"), ("vgg_q", (1, 3, 400, 533), "default"),
    ("vgg_f", (1, 3, 200, 200), "default"), ("vgg_q", (2, 3, 37, 53), "default"),
    ("vgg_q_he", (2, 3, 400, 400), "he"),
    ("resnet_h_he", (2, 3, 400, 400), "he"), ("resnet_f_he", (1, 3, 480, 640), "he")])
def test_network_forward_matches_oracle_fullres(name, shape, mode, built_lib):
    """BASELINE.json resolutions (400x400, shrink 533x400, 640x480) and an odd tiny size."""
    x = torch.rand(shape, generator=torch.Generator().manual_seed(5)) * 2 - 1
    sd = ref_models.synth_state_dict(_shapes(name), seed=1, mode=mode)
    gain = 1.0 / _oracle(name, sd, x[:1]).abs().max().item()
    sd = ref_models.synth_state_dict(_shapes(name), seed=1, out_gain=gain, mode=mode)
    net = _build(name, sd)
    ref = _oracle(name, sd, x).numpy()
    emu = _oracle(name, sd, x, emulate=True).numpy()
    with torch.no_grad():
        y = net(x.cuda())[0].cpu().numpy()
    assert y.shape == ref.shape
    err, err_emu = np.abs(y - ref).max(), np.abs(y - emu).max()
    print("%s %s [%s] max-abs vs fp32 oracle %.3g, vs fp16-operand emulation %.3g (ref max %.3g)"
          % (name, shape, mode, err, err_emu, np.abs(ref).max()))
    scale = max(1.0, np.abs(ref).max())
    floor = np.abs(emu - ref).max()
    assert err <= (BELIEF_TOL if mode == "default" else STRESS_TOL) * scale
    assert err <= FLOOR_FACTOR * floor + 1e-4 * scale, (err, floor)


@pytest.mark.parametrize("B,H,W,Cin,Cout,ksz,stride,extra", [
    (2, 32, 32, 64, 64, 3, 1, ""), (2, 40, 40, 64, 128, 3, 1, ""), (2, 25, 25, 128, 256, 3, 1, ""),
    (2, 50, 50, 256, 512, 3, 1, ""), (4, 25, 25, 512, 512, 3, 1, ""), (2, 100, 100, 64, 64, 3, 1, "res"),
    (2, 100, 100, 64, 7, 3, 1, "head"), (2, 50, 50, 256, 64, 1, 1, ""), (2, 50, 50, 128, 128, 3, 2, ""),
    (2, 25, 25, 256, 512, 1, 2, ""), (1, 13, 13, 2048, 256, 1, 1, ""),
    # row-shared kernel (conv_rs.cu): streamed weights N=128 / N=64, resident weights, partial tiles
    (2, 64, 64, 64, 128, 3, 1, ""), (2, 48, 64, 128, 128, 3, 1, ""), (2, 32, 32, 128, 64, 3, 1, ""),
    (1, 30, 45, 64, 64, 3, 1, ""), (2, 200, 200, 128, 128, 3, 1, "")])
def test_single_layers_are_exact_up_to_output_rounding(B, H, W, Cin, Cout, ksz, stride, extra, built_lib):
    """One conv layer on identical fp16 operands vs an fp32 conv: error <= one fp16 ulp of the output
    (the store rounding) + fp32 accumulation noise; the fp32 head output must agree to 1e-5 relative."""
    import torch.nn.functional as F
    from dream_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = (torch.randn((B, H, W, Cin), device="cuda", generator=g) * 0.5).half()
    w = torch.randn((Cout, Cin, ksz, ksz), device="cuda", generator=g) * (1.0 / (Cin * ksz * ksz) ** 0.5)
    b = torch.randn((Cout,), device="cuda", generator=g) * 0.1
    pad = ksz // 2
    Ho, Wo = (H + 2 * pad - ksz) // stride + 1, (W + 2 * pad - ksz) // stride + 1
    rs = [(r, s) for r in range(ksz) for s in range(ksz)]
    taps = [(r - pad, s - pad) for r, s in rs]
    ref = F.conv2d(x.permute(0, 3, 1, 2).double(), w.half().double(), b.double(), stride=stride, padding=pad).float()
    if extra == "head":
        wp = ops.pack_conv_weight(w, rs, cout_pad=16)
        got = ops.conv_taps(x, wp, ops.pad_bias(b, 16, "cuda"), taps, Ho, Wo, head_cout=Cout)
        assert (got - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()
        return
    wp = ops.pack_conv_weight(w, rs)
    residual = None
    if extra == "res":
        residual = torch.randn((B, Ho, Wo, wp.shape[1]), device="cuda", generator=g).half()
        ref = F.relu(ref + residual[..., :Cout].permute(0, 3, 1, 2).float())
    y = ops.conv_taps(x, wp, ops.pad_bias(b, wp.shape[1], "cuda"), taps, Ho, Wo, stride=stride,
                      relu=(extra == "res"), residual=residual)
    got = y[..., :Cout].permute(0, 3, 1, 2).float()
    tol = ref.abs() * 2.0 ** -11 + 1e-5 * ref.abs().max()
    assert bool(((got - ref).abs() <= tol).all()), ((got - ref).abs() - tol).max().item()
    if wp.shape[1] > Cout:
        assert float(y[..., Cout:].abs().max()) == 0.0          # padded channels stay exactly zero


@pytest.mark.parametrize("B,H,W,C,Co", [(2, 32, 48, 64, 64), (2, 50, 50, 128, 128), (1, 37, 53, 64, 128),
                                         (2, 100, 100, 256, 256), (3, 200, 200, 64, 64), (2, 64, 64, 128, 128),
                                         (1, 31, 46, 64, 64)])
def test_fused_pool_equals_conv_then_pool(B, H, W, C, Co, built_lib):
    """The 2x2 max pool fused into the conv epilogue is bit-identical to conv followed by the pool kernel."""
    from dream_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(7)
    x = (torch.randn((B, H, W, C), device="cuda", generator=g) * 0.5).half()
    w = torch.randn((Co, C, 3, 3), device="cuda", generator=g) * (1.0 / (C * 9) ** 0.5)
    b = torch.randn((Co,), device="cuda", generator=g) * 0.1
    rs = [(r, s) for r in range(3) for s in range(3)]
    wp, bp = ops.pack_conv_weight(w, rs), ops.pad_bias(b, ops.round_up(Co, 64), "cuda")
    y = ops.conv_taps(x, wp, bp, ops.TAPS_3x3, H, W, relu=True)
    ref_pool = ops.maxpool(y, 2, 2, 0)
    full, pooled = ops.conv_taps(x, wp, bp, ops.TAPS_3x3, H, W, relu=True, pool="both")
    assert torch.equal(full, y) and torch.equal(pooled, ref_pool)
    none, pooled2 = ops.conv_taps(x, wp, bp, ops.TAPS_3x3, H, W, relu=True, pool="only")
    assert none is None and torch.equal(pooled2, ref_pool)


@pytest.mark.parametrize("B,H,W", [(2, 64, 80), (1, 37, 53), (3, 400, 400)])
def test_fused_first_conv_matches_patch_path_and_torch(B, H, W, built_lib):
    """first_conv3x3 (gather + pack + MMA fused) vs the im2col + 1-tap GEMM path (same fp16 operands) and fp32."""
    import torch.nn.functional as F
    from dream_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(8)
    x = torch.rand((B, 3, H, W), device="cuda", generator=g) * 2 - 1
    w = torch.randn((64, 3, 3, 3), device="cuda", generator=g) * 0.2
    b = torch.randn((64,), device="cuda", generator=g) * 0.1
    wp, bp = ops.pack_first_weight(w, 64), ops.pad_bias(b, 64, "cuda")
    y = ops.first_conv3x3(x, wp, bp)
    y2 = ops.conv_taps(ops.im2col_first(x, 3, 3, 1, 1, 64), wp, bp, [(0, 0)], H, W, relu=True)
    ref = F.relu(F.conv2d(x.half().double(), w.half().double(), b.double(), padding=1)).float()
    got = y.permute(0, 3, 1, 2).float()
    tol = ref.abs() * 2.0 ** -11 + 1e-5 * ref.abs().max()
    assert bool(((got - ref).abs() <= tol).all())
    assert (y.float() - y2.float()).abs().max().item() <= 2.0 ** -10 * ref.abs().max().item()


def test_state_dict_round_trip_with_module_prefix(built_lib):
    from dream_b200 import models
    sd = ref_models.synth_state_dict(ref_models.vgg_state_shapes(7), seed=2)
    m = models.DataParallelShim(models.DreamHourglass(7, internalize_spatial_softmax=False))
    m.load_state_dict(sd, strict=True)
    out = m.state_dict()
    assert list(out.keys()) == list(sd.keys())
    for k in sd:
        assert torch.equal(out[k].cpu(), sd[k])
    sdr = ref_models.synth_state_dict(ref_models.resnet_state_shapes(7, full=True), seed=2)
    mr = models.DataParallelShim(models.ResnetSimple(7, full=True))
    mr.load_state_dict(sdr, strict=True)
    assert list(mr.state_dict().keys()) == list(sdr.keys())


@pytest.mark.parametrize("name,shape", [("resnet_h_he", None), ("resnet_f_he", None), ("resnet_h_he", (2, 3, 400, 400)),
                                        ("resnet_f_he", (1, 3, 480, 640))])
def test_resnet_precise_mode_meets_the_1e3_gate(name, shape, golden_dir, built_lib):
    """BASELINE's gate for configs 3 and 5: belief maps within 1e-3 max-abs of the fp32 reference.  The default path
    runs 11-bit tensor-core operands and sits at that format's floor (1.1e-3 .. 1.7e-3 on these He-scaled stress
    weights, STRESS_TOL above); `precise = True` (split-fp16 operands, same kernels, 3x the MMA work) must meet 1e-3
    with a wide margin -- on the reference-generated golden fixtures and at the BASELINE resolutions vs the oracle."""
    if shape is None:
        g = np.load(os.path.join(golden_dir, "net_%s.npz" % name))
        sd = ref_models.synth_state_dict(_shapes(name), seed=0, out_gain=float(g["gain"]), mode="he")
        x, ref = torch.from_numpy(g["x"]), g["y"]
    else:
        x = torch.rand(shape, generator=torch.Generator().manual_seed(5)) * 2 - 1
        sd = ref_models.synth_state_dict(_shapes(name), seed=1, mode="he")
        gain = 1.0 / _oracle(name, sd, x[:1]).abs().max().item()
        sd = ref_models.synth_state_dict(_shapes(name), seed=1, out_gain=gain, mode="he")
        ref = _oracle(name, sd, x).numpy()
    net = _build(name, sd)
    with torch.no_grad():
        fast = net(x.cuda())[0].cpu().numpy()
        net.precise = True
        y = net(x.cuda())[0].cpu().numpy()
    err, err_fast = np.abs(y - ref).max(), np.abs(fast - ref).max()
    print("%s %s precise max-abs %.3g (default path %.3g, ref max %.3g)" % (name, shape, err, err_fast, np.abs(ref).max()))
    assert y.shape == ref.shape
    assert err <= BELIEF_TOL * max(1.0, np.abs(ref).max()) / 5, err        # 2e-4 (measured 1.1e-4 at 480x640): five times inside the gate
    assert err < err_fast


def _flat(peaks):
    return np.array([(j, p[0], p[1], float(p[2]), p[3]) for j, lst in enumerate(peaks) for p in lst],
                    dtype=np.float64).reshape(-1, 5)


def test_peaks_kernel_matches_reference_golden(golden_dir, built_lib):
    from dream_b200 import image_proc
    g = np.load(os.path.join(golden_dir, "peaks.npz"))
    names = sorted({k.split("::")[0] for k in g.files})
    for name in names:
        maps = torch.from_numpy(g[name + "::maps"]).cuda()
        for off in (0.0, 0.4395):
            ref = g["%s::peaks@%g" % (name, off)]
            got = _flat(image_proc.peaks_from_belief_maps(maps, off))
            assert got.shape == ref.shape, (name, off, got.shape, ref.shape)
            assert np.array_equal(got[:, [0, 3, 4]], ref[:, [0, 3, 4]]), (name, off)
            assert np.array_equal(got[:, 1:3], ref[:, 1:3]), (name, off, np.abs(got[:, 1:3] - ref[:, 1:3]).max())


def test_peaks_kernel_integer_peaks_and_selection_match_oracle(built_lib):
    """Random smooth-ish maps at the BASELINE batch shape: integer peaks bit-exact, decisions identical."""
    from dream_b200 import image_proc
    rng = np.random.default_rng(11)
    B, K, H, W = 16, 7, 100, 100
    maps = np.zeros((B * K, H, W), np.float32)
    for i in range(B * K):
        pts = [(rng.uniform(0, W), rng.uniform(0, H)) for _ in range(int(rng.integers(0, 4)))]
        if pts:
            amp = rng.uniform(0.2, 1.0, size=len(pts))
            maps[i] = (ref_peaks.create_belief_map((W, H), pts) * amp[:, None, None]).sum(0)
        maps[i] += rng.standard_normal((H, W)).astype(np.float32) * 0.01
    table = image_proc.find_peaks_device(torch.from_numpy(maps).cuda(), 0.4395)
    sel = image_proc.select_keypoints_device(table, 0.25).cpu().numpy()
    counts = table.counts.cpu().numpy()
    ij = table.ij.cpu().numpy()
    ref = ref_peaks.peaks_from_belief_maps(maps, 0.4395)
    ref_sel = np.array(ref_peaks.select_keypoints(ref))
    for i in range(B * K):
        assert counts[i] == len(ref[i]), i
        sm = ref_peaks.gaussian_filter_f32(maps[i])
        ys, xs = np.nonzero(ref_peaks.peak_mask(sm))
        assert np.array_equal(ij[i, :counts[i], 0], xs) and np.array_equal(ij[i, :counts[i], 1], ys), i
    assert np.array_equal(sel, ref_sel)


def test_softargmax_kernel_matches_reference_golden(golden_dir, built_lib):
    from dream_b200.spatial_softmax import SoftArgmaxPavlo
    g = np.load(os.path.join(golden_dir, "softargmax.npz"))
    sa = SoftArgmaxPavlo(n_keypoints=7, learned_beta=True, initial_beta=25.0).cuda()
    with torch.no_grad():
        sa.beta.copy_(torch.from_numpy(g["beta"]))
        xy = sa(torch.from_numpy(g["maps"]).cuda()).cpu().numpy()
    assert np.abs(xy - g["xy"]).max() <= 1e-3


def test_facade_inference_end_to_end(built_lib):
    """DreamNetwork.inference: belief maps + keypoints vs oracle restatement of network.py:503-590."""
    from conftest import panda_config
    from dream_b200 import network
    net = network.create_network_from_config_data(panda_config("vgg"))
    assert net.trained_net_output_resolution() == (100, 100)
    sd = ref_models.synth_state_dict(ref_models.vgg_state_shapes(7), seed=3, out_gain=13.0, mode="default")
    net.model.load_state_dict(sd)
    net.enable_evaluation()
    x = torch.rand((2, 3, 400, 400), generator=torch.Generator().manual_seed(9)) * 2 - 1
    with torch.no_grad():
        belief, kps = net.inference(x.cuda())
    assert tuple(belief.shape) == (2, 7, 100, 100) and belief.is_cuda
    assert tuple(kps.shape) == (2, 7, 2) and not kps.is_cuda and kps.dtype == torch.float32
    ref_maps = ref_models.vgg_forward(sd, x).detach().numpy()
    assert np.abs(belief.cpu().numpy() - ref_maps).max() <= BELIEF_TOL * max(1.0, np.abs(ref_maps).max())
    # keypoint decisions on OUR maps must equal the reference algorithm applied to the same maps
    for b in range(2):
        ref_k = ref_peaks.select_keypoints(ref_peaks.peaks_from_belief_maps(belief[b].cpu().numpy(), 0.4395))
        assert np.array_equal(np.array(ref_k, dtype=np.float32), kps[b].numpy())


def test_conv_rejects_bad_arguments(built_lib):
    from dream_b200 import ops
    from dream_b200._lib import DreamB200Error
    x = torch.zeros((1, 8, 8, 64), dtype=torch.float16, device="cuda")
    w = torch.zeros((1, 32, 64), dtype=torch.float16, device="cuda")       # Cout_pad not a multiple of 64
    with pytest.raises(DreamB200Error):
        ops.conv_taps(x, w, None, [(0, 0)], 8, 8)


def test_missing_cpu_fallback_is_loud(built_lib):
    from dream_b200 import models
    net = models.DreamHourglass(7, internalize_spatial_softmax=False).eval()
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 3, 32, 32))


def test_prefetching_pipeline_matches_direct_inference(built_lib):
    """dream_b200.pipeline.inference_stream (H2D overlapped on a side stream) == DreamNetwork.inference."""
    from conftest import panda_config
    from dream_b200 import network, pipeline
    cfg = panda_config("vgg")
    cfg["training"]["config"]["net_input_resolution"] = [96, 64]
    net = network.create_network_from_config_data(cfg)
    sd = ref_models.synth_state_dict(ref_models.vgg_state_shapes(7), seed=4, out_gain=13.0, mode="default")
    net.model.load_state_dict(sd)
    net.enable_evaluation()
    gen = torch.Generator().manual_seed(2)
    batches = [(torch.rand((n, 3, 64, 96), generator=gen) * 2 - 1).pin_memory() for n in (3, 3, 3, 3, 2)]   # ragged end
    direct = []
    with torch.no_grad():
        for x in batches:
            b, k = net.inference(x.cuda())
            direct.append((b.cpu(), k))
    for graphs in (True, False, True):               # CUDA-graph pairs (captured, then reused) and eager launches
        net.use_cuda_graphs = graphs
        streamed = [(b.cpu(), k) for b, k in pipeline.inference_stream(net, batches)]
        assert len(streamed) == 5
        for (b0, k0), (b1, k1) in zip(direct, streamed):
            assert torch.equal(b0, b1) and torch.equal(k0, k1)
    assert len(net._stream_graphs) == 2              # one pair per batch shape, captured once


def test_facade_single_image_and_checkpoint_round_trip(tmp_path, built_lib):
    """keypoints_from_image (network.py:423-499) on a PIL frame, and save_network -> create_network_from_config_file
    (network.py:29-63, 592-632): identical keypoints after the round trip; 640x480 'shrink-and-crop' to 400x400."""
    from PIL import Image
    from conftest import panda_config
    from dream_b200 import network
    cfg = panda_config("vgg")
    cfg["training"]["config"]["net_input_resolution"] = [200, 200]
    net = network.create_network_from_config_data(cfg)
    sd = ref_models.synth_state_dict(ref_models.vgg_state_shapes(7), seed=8, out_gain=13.0, mode="default")
    net.model.load_state_dict(sd)
    net.enable_evaluation()
    rng = np.random.default_rng(0)
    img = Image.fromarray(rng.integers(0, 255, size=(480, 640, 3), dtype=np.uint8), "RGB")
    res = net.keypoints_from_image(img, debug=True)
    assert res["detected_keypoints"].shape == (7, 2)
    assert res["image_rgb_net_input"].size == (200, 200)
    assert tuple(res["belief_maps"].shape) == (7, 48, 48)      # 200 -> 100 -> 50 -> 25 -> 12 (floor) -> x4
    # the same frame through the oracle: preprocess -> normalise -> reference forward -> reference peak logic
    arr = np.asarray(res["image_rgb_net_input"].convert("RGB"), dtype=np.float32) / 255.0
    x = torch.from_numpy(((arr - 0.5) / 0.5).transpose(2, 0, 1)[None].copy())
    ref_maps = ref_models.vgg_forward(sd, x).detach().numpy()[0]
    assert np.abs(res["belief_maps"].cpu().numpy() - ref_maps).max() <= BELIEF_TOL * max(1.0, np.abs(ref_maps).max())
    ref_k = np.array(ref_peaks.select_keypoints(
        ref_peaks.peaks_from_belief_maps(res["belief_maps"].cpu().numpy(), 0.4395)), dtype=np.float32)
    assert np.array_equal(ref_k, res["detected_keypoints_net_output"].astype(np.float32))
    # checkpoint round trip
    net.save_network(str(tmp_path), "ckpt", overwrite=True)
    assert (tmp_path / "ckpt.pth").exists() and (tmp_path / "ckpt.yaml").exists()
    assert all(k.startswith("module.") for k in torch.load(tmp_path / "ckpt.pth").keys())
    net2 = network.create_network_from_config_file(str(tmp_path / "ckpt.yaml"), str(tmp_path / "ckpt.pth"))
    net2.enable_evaluation()
    res2 = net2.keypoints_from_image(img)
    assert np.array_equal(res["detected_keypoints"], res2["detected_keypoints"])


@pytest.mark.gpu
def test_pair_kernel_is_bit_identical_to_single_cta_kernel(built_lib):
    """conv_rs2 (tcgen05 cta_group::2 CTA pair, conv_rs2.cu) against conv_rs on the same inputs: same MMA K order,
    same epilogue, so the outputs must agree bit for bit (64 -> 64 with / without the fused pool, odd tile counts,
    128 output channels through the opt-in mask).  Each side runs in its own process (tools/rs2_check.py): the
    kernel choice is read once per process from DREAMB200_RS2."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "rs2_check.py"), "s64_64_40x56_pool",
                        "s64_64_100_both", "s128_128_48_pool", "g128_128_64_bwd"], capture_output=True, text=True, timeout=1500)
    assert "ALL IDENTICAL" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]

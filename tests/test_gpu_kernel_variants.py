"""GPU tests (-m gpu) of the kernels that have a second implementation of the same op: each default kernel must agree
with its fallback through the C-ABI -- bit for bit where the arithmetic order is the same (fused vs three-kernel peak
extraction, conv_tc2 vs conv_tc), to fp32 atomic-order noise where it is not (wgrad3x3_pair vs wgrad3x3) -- and with
the oracle.  The switches are environment variables the library reads on every call."""
import os
from contextlib import contextmanager

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ref_peaks  # noqa: E402  (checker only)


@contextmanager
def env(**kw):
    old = {k: os.environ.get(k) for k in kw}
    try:
        for k, v in kw.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        yield
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _table_np(t):
    return {k: getattr(t, k).cpu().numpy() for k in ("xy", "score", "ij", "counts", "summary")}


def _maps(rng, n, h, w, kind):
    if kind == "blobs":
        m = np.zeros((n, h, w), np.float32)
        yy, xx = np.mgrid[0:h, 0:w]
        for i in range(n):
            for _ in range(int(rng.integers(0, 5))):
                cx, cy, a = rng.uniform(-2, w + 2), rng.uniform(-2, h + 2), rng.uniform(0.05, 1.0)
                m[i] += (a * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / 8.0)).astype(np.float32)
            m[i] += rng.standard_normal((h, w)).astype(np.float32) * 0.01
        return m
    if kind == "noise":                     # many local maxima above the threshold: exercises the ordered compaction
        return (rng.standard_normal((n, h, w)) * 0.5 + 0.3).astype(np.float32)
    if kind == "plateau":                   # every pixel passes the >= test: counts far above the table capacity
        m = np.full((n, h, w), 0.5, np.float32)
        m[1::2] = 0.0                       # ... and all-zero maps: no peak at all
        return m
    raise ValueError(kind)


@pytest.mark.parametrize("h,w,kind", [(100, 100, "blobs"), (100, 100, "noise"), (37, 53, "blobs"), (25, 30, "noise"), (5, 7, "noise"),
                                      (1, 9, "noise"), (120, 160, "blobs"), (64, 48, "plateau"), (13, 11, "plateau")])
def test_fused_peaks_equal_three_kernel_path_and_oracle(h, w, kind, built_lib):
    from dream_b200 import image_proc
    rng = np.random.default_rng(h * 1000 + w)
    n = 21
    maps = _maps(rng, n, h, w, kind)
    dev = torch.from_numpy(maps).cuda()
    with env(DREAMB200_PEAKS_UNFUSED=None):
        fused = _table_np(image_proc.find_peaks_device(dev, 0.4395, cap=32))
    with env(DREAMB200_PEAKS_UNFUSED="1"):
        three = _table_np(image_proc.find_peaks_device(dev, 0.4395, cap=32))
    assert np.array_equal(fused["counts"], three["counts"])
    for i in range(n):
        c = min(int(fused["counts"][i]), 32)
        for k in ("xy", "score", "ij"):
            assert np.array_equal(fused[k][i, :c], three[k][i, :c]), (i, k)
        if fused["counts"][i] > 0:
            assert np.array_equal(fused["summary"][i], three["summary"][i]), i
    # ... and the integer peak set equals scipy's (oracle) on every map
    for i in range(n):
        sm = ref_peaks.gaussian_filter_f32(maps[i])
        ys, xs = np.nonzero(ref_peaks.peak_mask(sm))
        assert fused["counts"][i] == len(xs), i
        c = min(len(xs), 32)
        assert np.array_equal(fused["ij"][i, :c, 0], xs[:c]) and np.array_equal(fused["ij"][i, :c, 1], ys[:c]), i


@pytest.mark.parametrize("h,w,n,kind", [(100, 100, 21, "blobs"), (37, 53, 5, "noise"), (208, 208, 14, "blobs"),
                                        (400, 400, 7, "blobs"), (480, 640, 7, "blobs"), (208, 208, 3, "noise"),
                                        (300, 40, 4, "plateau"), (26, 700, 2, "noise")])
def test_banded_peaks_equal_the_other_paths_and_oracle(h, w, n, kind, built_lib):
    """peaks_banded_kernel (one CTA per band of rows; resnet-H 208x208 and the full-resolution decoders' maps) against
    the three generic kernels -- and the whole-map kernel where the map fits -- on the same maps: identical tables,
    counts and summaries; integer peak set equal to scipy's (oracle) on a sample.  Small batches make the bands short
    (more CTAs than SMs), so band seams cross peaks, plateaus and the reflected image border."""
    from dream_b200 import image_proc
    rng = np.random.default_rng(h * 7 + w)
    maps = _maps(rng, n, h, w, kind)
    dev = torch.from_numpy(maps).cuda()
    cap = 48
    with env(DREAMB200_PEAKS_BANDED="1", DREAMB200_PEAKS_UNFUSED=None):
        banded = _table_np(image_proc.find_peaks_device(dev, 0.4395, cap=cap))
    with env(DREAMB200_PEAKS_BANDED=None, DREAMB200_PEAKS_UNFUSED="1"):
        three = _table_np(image_proc.find_peaks_device(dev, 0.4395, cap=cap))
    with env(DREAMB200_PEAKS_BANDED=None, DREAMB200_PEAKS_UNFUSED=None):
        default = _table_np(image_proc.find_peaks_device(dev, 0.4395, cap=cap))
    for other in (three, default):
        assert np.array_equal(banded["counts"], other["counts"])
        for i in range(n):
            c = min(int(banded["counts"][i]), cap)
            for k in ("xy", "score", "ij"):
                assert np.array_equal(banded[k][i, :c], other[k][i, :c]), (i, k)
            if banded["counts"][i] > 0:
                assert np.array_equal(banded["summary"][i], other["summary"][i]), i
    for i in range(min(n, 3)):
        ys, xs = np.nonzero(ref_peaks.peak_mask(ref_peaks.gaussian_filter_f32(maps[i])))
        assert banded["counts"][i] == len(xs), i
        c = min(len(xs), cap)
        assert np.array_equal(banded["ij"][i, :c, 0], xs[:c]) and np.array_equal(banded["ij"][i, :c, 1], ys[:c]), i


@pytest.mark.parametrize("h,w", [(100, 100), (37, 53), (7, 5), (120, 160)])
def test_gaussian_smoothing_is_bit_exact_including_edge_values(h, w, built_lib):
    """The smoothing stage alone against the oracle's scipy restatement, bit for bit, on values that exercise every
    branch of the integer-pipe fp32<->fp64 conversions: normal numbers of both signs and all magnitudes, exact zeros,
    fp32 subnormals (inputs and results), values near FLT_MAX."""
    from dream_b200 import image_proc
    rng = np.random.default_rng(h + w)
    n = 12
    maps = rng.standard_normal((n, h, w)).astype(np.float32)
    maps[1] *= np.float32(1e-30)
    maps[2] = (rng.standard_normal((h, w)) * 1e-40).astype(np.float32)              # subnormal inputs
    maps[3] = 0.0
    maps[4] = np.where(rng.random((h, w)) < 0.7, 0.0, maps[4])                      # mostly exact zeros
    maps[5] *= np.float32(1e30)
    maps[6] = np.float32(3e38) * np.sign(maps[6])                                   # near FLT_MAX, both signs
    maps[7] = (10.0 ** rng.uniform(-44, 38, size=(h, w)) * np.sign(maps[7])).astype(np.float32)
    maps[8] = np.float32(1.17549435e-38) * np.where(rng.random((h, w)) < 0.5, 1, -1)  # smallest normal: subnormal sums
    maps[9, ::2] = 0.0
    maps[9, 1::2] = np.float32(1e-45)                                               # smallest subnormal
    want = np.stack([ref_peaks.gaussian_filter_f32(m) for m in maps])
    dev = torch.from_numpy(maps).cuda()
    for unfused in (None, "1"):
        with env(DREAMB200_PEAKS_UNFUSED=unfused):
            got = image_proc.gaussian_smooth_device(dev).cpu().numpy()
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), \
            (unfused, np.argwhere(got.view(np.uint32) != want.view(np.uint32))[:5])


def test_fused_peaks_full_batch_shape_matches_oracle_decisions(built_lib):
    """B=128 x 7 maps of 100x100 (the benchmarked shape) through the fused kernel: counts, refined coordinates and the
    keypoint decision equal the oracle's on a sample of the maps, and the two device paths agree on all of them."""
    from dream_b200 import image_proc
    rng = np.random.default_rng(5)
    maps = _maps(rng, 896, 100, 100, "blobs")
    dev = torch.from_numpy(maps).cuda()
    t = image_proc.find_peaks_device(dev, 0.4395)
    sel = image_proc.select_keypoints_device(t, 0.25).cpu().numpy()
    with env(DREAMB200_PEAKS_UNFUSED="1"):
        t3 = image_proc.find_peaks_device(dev, 0.4395)
        sel3 = image_proc.select_keypoints_device(t3, 0.25).cpu().numpy()
    assert np.array_equal(sel, sel3) and torch.equal(t.counts, t3.counts)
    idx = rng.choice(896, 48, replace=False)
    ref = ref_peaks.peaks_from_belief_maps(maps[idx], 0.4395)
    ref_sel = np.array(ref_peaks.select_keypoints(ref))
    assert np.array_equal(sel[idx], ref_sel)
    xy = t.xy.cpu().numpy()
    for j, i in enumerate(idx):
        assert int(t.counts[i]) == len(ref[j])
        for s, p in enumerate(ref[j][:64]):
            assert xy[i, s, 0] == p[0] and xy[i, s, 1] == p[1]


CONV_CASES = {   # name: (B, H, W, Cin, Cout, kind)
    "256_256_50": (2, 50, 50, 256, 256, "3x3"),
    "128_256_100_pool": (2, 100, 100, 128, 256, "pool"),
    "256_512_26x30_s2": (2, 26, 30, 256, 512, "s2"),
    "256_1024_25_1x1res": (2, 25, 25, 256, 1024, "1x1res"),
    "64_256_100_1x1res": (3, 100, 100, 64, 256, "1x1res"),
    "512_2048_13_1x1res": (5, 13, 13, 512, 2048, "1x1res"),
    "512_512_25_odd_tiles": (1, 25, 25, 512, 512, "3x3"),
}


@pytest.mark.parametrize("name", sorted(CONV_CASES))
def test_conv_pair_kernel_is_bit_identical_to_single_cta(name, built_lib):
    from dream_b200 import ops
    B, H, W, Cin, Cout, kind = CONV_CASES[name]
    g = torch.Generator(device="cuda").manual_seed(11)
    x = (torch.randn((B, H, W, Cin), device="cuda", generator=g) * 0.5).half()
    ksz = 1 if kind == "1x1res" else 3
    stride = 2 if kind == "s2" else 1
    pad = ksz // 2
    w = torch.randn((Cout, Cin, ksz, ksz), device="cuda", generator=g) * (1.0 / (Cin * ksz * ksz) ** 0.5)
    bias = torch.randn((Cout,), device="cuda", generator=g) * 0.1
    rs = [(r, s) for r in range(ksz) for s in range(ksz)]
    taps = [(r - pad, s - pad) for r, s in rs]
    Ho, Wo = (H + 2 * pad - ksz) // stride + 1, (W + 2 * pad - ksz) // stride + 1
    wp, bp = ops.pack_conv_weight(w, rs), ops.pad_bias(bias, Cout, "cuda")
    res32 = torch.randn((B, Ho, Wo, Cout), device="cuda", generator=g) if kind == "1x1res" else None

    def run():
        kw = {"relu": True, "stride": stride}
        if kind == "pool":
            kw["pool"] = "both"
        if kind == "1x1res":
            kw["residual_f32"] = res32
            kw["y_f32"] = torch.empty((B, Ho, Wo, Cout), device="cuda")
        y = ops.conv_taps(x, wp, bp, taps, Ho, Wo, **kw)
        outs = [t for t in (y if isinstance(y, tuple) else (y,)) if t is not None]
        if kind == "1x1res":
            outs.append(kw["y_f32"])
        torch.cuda.synchronize()
        return outs

    with env(DREAMB200_TC2="0"):
        single = run()
    with env(DREAMB200_TC2="1", DREAMB200_RES_TMA="0"):
        pair = run()
    with env(DREAMB200_TC2="1", DREAMB200_RES_TMA="1", DREAMB200_RES_INPLACE="0"):   # fp32 residual staged by TMA
        pair_tma = run()
    with env(DREAMB200_TC2="1", DREAMB200_RES_TMA="1", DREAMB200_RES_INPLACE="1"):   # ... fp32 output through the same slots (default)
        pair_inplace = run()
    assert len(single) == len(pair) == len(pair_tma) == len(pair_inplace)
    for a, b, c, d in zip(single, pair, pair_tma, pair_inplace):
        assert torch.equal(a, b) and torch.equal(a, c) and torch.equal(a, d)
    # ... and both are the convolution (fp32 torch reference of the same fp16 operands)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).float(), w.half().float(), bias, stride=stride, padding=pad)
    if res32 is not None:
        ref = ref + res32.permute(0, 3, 1, 2)
    ref = torch.relu(ref).permute(0, 2, 3, 1)
    got = pair[-1] if kind == "1x1res" else pair[0]
    assert (got.float() - ref).abs().max().item() <= 4e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("B,H,W", [(2, 64, 48), (1, 100, 100), (3, 37, 53), (2, 33, 41), (1, 400, 400), (2, 32, 8),
                                   (1, 96, 17)])
@pytest.mark.parametrize("pool", [None, "only"])
def test_two_row_slab_kernel_is_bit_identical_to_the_pair_kernel(B, H, W, pool, built_lib):
    """conv_rs3 (one accumulator row = two vertically adjacent output pixels, 64 -> 64 channels) against conv_rs2 /
    conv_rs on the same layer: every output receives the same partial products in the same order -> identical bits,
    with and without the fused 2x2 max pool, on full, ragged and odd-sized maps; and both are the convolution."""
    from dream_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(H * 1000 + W)
    x = (torch.randn((B, H, W, 64), device="cuda", generator=g) * 0.5).half()
    w = torch.randn((64, 64, 3, 3), device="cuda", generator=g) * (1.0 / (64 * 9) ** 0.5)
    bias = torch.randn((64,), device="cuda", generator=g) * 0.1
    rs = [(r, s) for r in range(3) for s in range(3)]
    wp, bp = ops.pack_conv_weight(w, rs), ops.pad_bias(bias, 64, "cuda")

    def run():
        y = ops.conv_taps(x, wp, bp, ops.TAPS_3x3, H, W, relu=True, pool=pool)
        outs = [t for t in (y if isinstance(y, tuple) else (y,)) if t is not None]
        torch.cuda.synchronize()
        return outs

    with env(DREAMB200_RS3="0"):
        old = run()
    with env(DREAMB200_RS3="3"):
        new = run()
    assert len(old) == len(new) == 1
    assert torch.equal(old[0], new[0])
    torch.backends.cudnn.allow_tf32 = False
    ref = torch.relu(torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).float(), w.half().float(), bias, padding=1))
    if pool is not None:
        ref = torch.nn.functional.max_pool2d(ref, 2)
    assert (new[0].float() - ref.permute(0, 2, 3, 1)).abs().max().item() <= 4e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("B,H,W,Ci,Co", [(3, 25, 25, 128, 256), (2, 50, 37, 256, 256), (2, 13, 13, 512, 512),
                                         (2, 1, 9, 256, 512), (1, 100, 100, 128, 256), (2, 26, 25, 128, 256),
                                         (1, 27, 16, 128, 256)])
def test_wgrad_pair_kernel_matches_single_cta_and_autograd(B, H, W, Ci, Co, built_lib):
    from dream_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    x = (torch.randn((B, H, W, Ci), device="cuda", generator=g) * 0.5).half()
    dy = (torch.randn((B, H, W, Co), device="cuda", generator=g) * 0.5).half()
    with env(DREAMB200_WGRAD3_2SM="0"):
        single = ops.wgrad(dy, x, ops.TAPS_3x3)
    with env(DREAMB200_WGRAD3_2SM="1"):
        pair = ops.wgrad(dy, x, ops.TAPS_3x3)
    torch.cuda.synchronize()
    scale = single.abs().max().item()
    assert (single - pair).abs().max().item() <= 1e-5 * scale          # same products, different fp32 summation order
    xr = x.permute(0, 3, 1, 2).float()
    wz = torch.zeros((Co, Ci, 3, 3), device="cuda", requires_grad=True)
    torch.backends.cudnn.allow_tf32 = False
    torch.nn.functional.conv2d(xr, wz, padding=1).backward(dy.permute(0, 3, 1, 2).float())
    ref = wz.grad.permute(2, 3, 0, 1).reshape(9, Co, Ci)
    assert (pair - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
    # the pair kernel picks its tile height from the map (10 rows at 50x50, 20 at 100x100, 26 at 25x25 ...): every
    # pinned height must give the same sums
    for th in ("16", "6", "22"):
        with env(DREAMB200_WGRAD3_2SM="1", DREAMB200_WGRAD_TH=th):
            pinned = ops.wgrad(dy, x, ops.TAPS_3x3)
        assert (single - pinned).abs().max().item() <= 1e-5 * scale, th


def test_cuda_graph_inference_is_bit_identical_to_eager_and_follows_the_weights(built_lib):
    """DreamNetwork.capture_inference / inference_graphed (the whole step as one CUDA graph) against the eager launches:
    same kernels on the same data -> identical bits; new weights or a new shape -> a new capture, not a stale replay."""
    from conftest import panda_config
    from dream_b200 import network
    from oracle import ref_models
    net = network.create_network_from_config_data(panda_config("vgg"))
    net.enable_evaluation()
    sd = ref_models.synth_state_dict(ref_models.vgg_state_shapes(7), seed=3, out_gain=13.0, mode="default")
    net.model.load_state_dict(sd)
    g = torch.Generator().manual_seed(1)
    xa = (torch.rand((3, 3, 96, 128), generator=g) * 2 - 1).cuda()
    xb = (torch.rand((3, 3, 96, 128), generator=g) * 2 - 1).cuda()
    with torch.no_grad():
        for x in (xa, xb, xa):
            be, ke = net.inference_device(x)
            bg, kg = net.inference_graphed(x)
            assert torch.equal(be, bg) and torch.equal(ke, kg)
        assert len(net._graph_cache) == 1
        graph = next(iter(net._graph_cache.values()))
        assert graph.replays == 3 and graph.kernels_per_replay >= 20
        # explicit capture on the caller's buffer: refill in place, replay
        buf = xa.clone()
        cap = net.capture_inference(buf, adopt=True)
        buf.copy_(xb)
        b2, k2 = cap()
        be, ke = net.inference_device(xb)
        assert torch.equal(b2, be) and torch.equal(k2, ke)
        # another shape -> another graph; new weights -> the old graph is not reused
        xc = (torch.rand((1, 3, 64, 64), generator=g) * 2 - 1).cuda()
        assert torch.equal(net.inference_graphed(xc)[0], net.inference_device(xc)[0])
        sd2 = ref_models.synth_state_dict(ref_models.vgg_state_shapes(7), seed=4, out_gain=13.0, mode="default")
        net.model.load_state_dict(sd2)
        be, ke = net.inference_device(xa)
        bg, kg = net.inference_graphed(xa)
        assert torch.equal(be, bg) and torch.equal(ke, kg) and not torch.equal(be, b2)
    # the single-image facade goes through the graph by default and must agree with the eager facade
    from PIL import Image
    img = Image.fromarray((np.random.default_rng(0).random((480, 640, 3)) * 255).astype(np.uint8))
    ra = net.keypoints_from_image(img)
    net.use_cuda_graphs = False
    rb = net.keypoints_from_image(img)
    assert np.array_equal(ra["detected_keypoints"], rb["detected_keypoints"])


@pytest.mark.parametrize("kind,shape", [("vgg_q", (3, 3, 200, 200)), ("vgg_q", (2, 3, 104, 72)), ("vgg_f", (2, 3, 96, 128)),
                                        ("resnet_h", (2, 3, 160, 192)), ("resnet_f", (1, 3, 96, 128))])
def test_phase_groups_are_bit_identical_to_phase_by_phase_launches(kind, shape, built_lib):
    """dreamb200_conv2d_fwd_phases: the four sub-pixel phases of a folded upsample + conv (vgg-Q: 512 -> 256 on the
    CTA-pair N = 256 kernel, 256 -> 128 on its N = 128 form), of ConvTranspose(4,2,1) (resnet, 256 channels) and of
    ConvTranspose(3,2,1,1) (vgg-F: 1/2/2/4 taps -- NOT uniform, must fall back) as one launch vs one launch per phase
    vs the single-CTA kernel: the belief maps must not change by a bit."""
    from dream_b200 import models
    from oracle import ref_models
    if kind.startswith("vgg"):
        kw = dict(deconv_decoder=True, full_output=True) if kind == "vgg_f" else {}
        sd = ref_models.synth_state_dict(ref_models.vgg_state_shapes(7, prefix="", **kw), seed=2, out_gain=13.0, mode="default")
        net = models.DreamHourglass(7, internalize_spatial_softmax=False, **kw)
    else:
        full = kind == "resnet_f"
        sd = ref_models.synth_state_dict(ref_models.resnet_state_shapes(7, full=full, prefix=""), seed=2, out_gain=0.05, mode="he")
        net = models.ResnetSimple(7, full=full, pretrained=False)
    net.load_state_dict(sd)
    net = net.cuda().eval()
    x = (torch.rand(shape, generator=torch.Generator().manual_seed(0)) * 2 - 1).cuda()
    from dream_b200 import _lib
    outs, launches = [], []
    with torch.no_grad():
        net.belief_maps(x)                           # packs the weights
        for groups, tc2 in (("1", "1"), ("0", "1"), ("0", "0")):
            with env(DREAMB200_PHASE_GROUPS=groups, DREAMB200_TC2=tc2):
                n0 = _lib.launch_count()
                outs.append(net.belief_maps(x).clone())
                launches.append(_lib.launch_count() - n0)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    if kind != "vgg_f" and shape[2] * shape[3] >= 160 * 160:
        assert launches[0] < launches[1], launches   # the groups really ran as fewer launches (tiny maps: < 2 M-tiles)


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 64, 48, 64, 64), (3, 400, 400, 64, 64), (2, 200, 200, 128, 128),
                                            (2, 100, 100, 256, 256), (1, 50, 52, 128, 256), (2, 34, 22, 64, 128)])
def test_pool_first_epilogue_is_bit_identical(B, H, W, Cin, Cout, built_lib):
    """Pool-only launches (inference) take the epilogue that pools the fp32 accumulators first and finishes only the
    pooled quarter; it must equal, bit for bit, the pooled tensor of a launch that also stores the full tile (which
    pools the finished fp16 values), and max_pool2d of that full tile."""
    from dream_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(7)
    x = (torch.randn((B, H, W, Cin), device="cuda", generator=g) * 0.5).half()
    w = torch.randn((Cout, Cin, 3, 3), device="cuda", generator=g) * (1.0 / (Cin * 9) ** 0.5)
    bias = torch.randn((Cout,), device="cuda", generator=g) * 0.3
    rs = [(r, s) for r in range(3) for s in range(3)]
    wp, bp = ops.pack_conv_weight(w, rs), ops.pad_bias(bias, Cout, "cuda")
    for relu in (True, False):
        full, pooled_both = ops.conv_taps(x, wp, bp, ops.TAPS_3x3, H, W, relu=relu, pool="both")
        _, pooled_only = ops.conv_taps(x, wp, bp, ops.TAPS_3x3, H, W, relu=relu, pool="only")
        torch.cuda.synchronize()
        assert torch.equal(pooled_both, pooled_only)
        ref = torch.nn.functional.max_pool2d(full.permute(0, 3, 1, 2).float(), 2).permute(0, 2, 3, 1).half()
        assert torch.equal(pooled_only, ref)

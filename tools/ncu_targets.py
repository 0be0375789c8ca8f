"""Single-kernel targets for `ncu --set full`: python tools/ncu_targets.py <target> [batch].
Each target launches its kernel 3 times on BASELINE-shaped tensors (profile the 3rd: --launch-skip 2 -c 1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dream_b200 import ops

target = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
g = torch.Generator(device="cuda").manual_seed(0)


def conv_case(H, W, Cin, Cout, pool=None):
    x = (torch.randn((B, H, W, Cin), device="cuda", generator=g) * 0.5).half()
    w = torch.randn((Cout, Cin, 3, 3), device="cuda", generator=g) * (1.0 / (Cin * 9) ** 0.5)
    rs = [(r, s) for r in range(3) for s in range(3)]
    wp, bp = ops.pack_conv_weight(w, rs), ops.pad_bias(None, ops.round_up(Cout, 64), "cuda")
    for _ in range(3):
        ops.conv_taps(x, wp, bp, ops.TAPS_3x3, H, W, relu=True, pool=pool)


if target == "conv256":            # layer_0_3_down.12/.14: 256->256 @100x100 (conv_tc_kernel<256,0>), 5.9 GMAC/img
    conv_case(100, 100, 256, 256)
elif target == "conv512":          # layer_0_4_down.21..25: 512->512 @50x50
    conv_case(50, 50, 512, 512)
elif target == "rs64":             # layer_0_1_down.2: 64->64 @400x400 + fused pool (conv_rs3_kernel<true> since round 2)
    conv_case(400, 400, 64, 64, pool="only")
elif target == "dgrad64":         # data gradient of layer_0_1_down.2 (64 -> 64 @400x400): ReLU gate staged by TMA, re-scaling,
    # bias-gradient column sums, running max (conv_rs2_kernel<64, ., false> with the gate ring)
    dy = (torch.randn((B, 400, 400, 64), device="cuda", generator=g) * 0.5).half()
    w = torch.randn((64, 64, 3, 3), device="cuda", generator=g) * (1.0 / (64 * 9) ** 0.5)
    wp = ops.pack_conv_weight(w, [(r, s) for r in range(3) for s in range(3)])
    gate = torch.relu(torch.randn((B, 400, 400, 64), device="cuda", generator=g)).half()
    scale = torch.full((1,), 2.0, device="cuda")
    for _ in range(3):
        amax = torch.zeros((1,), device="cuda"); col = torch.zeros((64,), device="cuda")
        ops.conv_taps(dy, wp, None, ops.TAPS_3x3, 400, 400, gate=gate, out_scale=scale, colsum=col, absmax=amax)
elif target == "rs128":            # layer_0_2_down.7: 128->128 @200x200 + fused pool (conv_rs_kernel<128,false>)
    conv_case(200, 200, 128, 128, pool="only")
elif target == "first":            # layer_0_1_down.0 fused with the input pack (first_conv_kernel)
    x = torch.rand((B, 3, 400, 400), device="cuda", generator=g) * 2 - 1
    w = torch.randn((64, 3, 3, 3), device="cuda", generator=g) * 0.2
    wp, bp = ops.pack_first_weight(w, 64), ops.pad_bias(None, 64, "cuda")
    for _ in range(3):
        ops.first_conv3x3(x, wp, bp)
elif target == "wgrad256":
    x = (torch.randn((B, 100, 100, 256), device="cuda", generator=g) * 0.5).half()
    dy = (torch.randn((B, 100, 100, 256), device="cuda", generator=g) * 0.5).half()
    for _ in range(3):
        ops.wgrad(dy, x, ops.TAPS_3x3)
elif target == "wgrad512":         # 512 -> 512 @50x50: wgrad3x3_pair_kernel with 10-row tiles (round 2: 1.23 -> 1.03 ms)
    x = (torch.randn((B, 50, 50, 512), device="cuda", generator=g) * 0.5).half()
    dy = (torch.randn((B, 50, 50, 512), device="cuda", generator=g) * 0.5).half()
    for _ in range(3):
        ops.wgrad(dy, x, ops.TAPS_3x3)
elif target == "expand":           # resnet layer3 conv3: 1x1 256 -> 1024 @25x25 + fp32 identity in / out (conv_tc2, residual ring)
    Bx = 64
    x = (torch.randn((Bx, 25, 25, 256), device="cuda", generator=g) * 0.5).half()
    w = torch.randn((1024, 256, 1, 1), device="cuda", generator=g) * (1.0 / 256 ** 0.5)
    wp, bp = ops.pack_conv_weight(w, [(0, 0)]), ops.pad_bias(None, 1024, "cuda")
    res = torch.randn((Bx, 25, 25, 1024), device="cuda", generator=g)
    y32 = torch.empty_like(res)
    for _ in range(3):
        ops.conv_taps(x, wp, bp, [(0, 0)], 25, 25, relu=True, residual_f32=res, y_f32=y32)
elif target == "peaks208":         # resnet-H shape: 64 x 7 maps of 208x208 (peaks_banded_kernel)
    from dream_b200 import image_proc
    maps = torch.randn((64 * 7, 208, 208), device="cuda", generator=g) * 0.2
    maps[:, 90:94, 120:124] += 1.0
    for _ in range(3):
        image_proc.find_peaks_device(maps, 0.4395)
elif target == "peaks":            # peak extraction on B*7 belief maps of 100x100 (peaks_fused_kernel)
    from dream_b200 import image_proc
    maps = torch.randn((B * 7, 100, 100), device="cuda", generator=g) * 0.2
    maps[:, 40:44, 60:64] += 1.0
    for _ in range(3):
        image_proc.find_peaks_device(maps, 0.4395)
torch.cuda.synchronize()

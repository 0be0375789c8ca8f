"""GPU bring-up harness for the tcgen05 conv kernel: each case runs in its own subprocess
(a trapped kernel kills only that case) and is compared against torch conv2d in fp32 on
fp16-rounded operands.  Writes gpurun_out/conv_check.json."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [
    # name, B, H, W, Cin, Cout, mode
    ("c64_64_16x8", 1, 8, 16, 64, 64, "3x3"),
    ("c64_64_32x32", 2, 32, 32, 64, 64, "3x3"),
    ("c64_128_40x40", 2, 40, 40, 64, 128, "3x3"),
    ("c128_256_25x25", 2, 25, 25, 128, 256, "3x3"),
    ("c256_512_50x50", 2, 50, 50, 256, 512, "3x3"),
    ("c512_512_25x25_b4", 4, 25, 25, 512, 512, "3x3"),
    ("c64_64_100x100_relu_res", 2, 100, 100, 64, 64, "3x3_res"),
    ("head_64_7_100x100", 2, 100, 100, 64, 7, "head"),
    ("1x1_256_64_50x50", 2, 50, 50, 256, 64, "1x1"),
    ("first_3x3_64_80x60", 2, 60, 80, 3, 64, "first3"),
    ("first_7x7s2_64_80x60", 2, 60, 80, 3, 64, "first7"),
    ("s2_3x3_128_128_50x50", 2, 50, 50, 128, 128, "3x3s2"),
    ("s2_1x1_256_512_25x25", 2, 25, 25, 256, 512, "1x1s2"),
    ("pool_up", 2, 26, 30, 64, 64, "poolup"),
    ("big_c64_64_400x400_b8", 8, 400, 400, 64, 64, "3x3"),
    ("big_c256_256_100x100_b32", 32, 100, 100, 256, 256, "3x3"),
]


def run_case(name):
    import torch
    import torch.nn.functional as F
    from dream_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    case = [c for c in CASES if c[0] == name][0]
    _, B, H, W, Cin, Cout, mode = case
    g = torch.Generator(device="cuda").manual_seed(1234)
    dev = "cuda"
    res = {"name": name}
    if mode in ("first3", "first7"):
        R = 3 if mode == "first3" else 7
        stride = 1 if mode == "first3" else 2
        pad = R // 2
        x = torch.rand((B, 3, H, W), device=dev, generator=g) * 2 - 1
        w = torch.randn((Cout, 3, R, R), device=dev, generator=g) * 0.1
        b = torch.randn((Cout,), device=dev, generator=g) * 0.1
        Kpad = ops.round_up(R * R * 3, 64)
        patches = ops.im2col_first(x, R, R, stride, pad, Kpad)
        wp = ops.pack_first_weight(w, Kpad)
        Ho, Wo = patches.shape[1], patches.shape[2]
        y = ops.conv_taps(patches, wp, ops.pad_bias(b, wp.shape[1], dev), [(0, 0)], Ho, Wo, relu=True)
        ref = F.relu(F.conv2d(x.half().float(), w.half().float(), b, stride=stride, padding=pad))
        got = y[..., :Cout].permute(0, 3, 1, 2).float()
    elif mode == "poolup":
        x = torch.randn((B, H, W, Cin), device=dev, generator=g).half()
        p2 = ops.maxpool(x, 2, 2, 0)
        ref_p2 = F.max_pool2d(x.permute(0, 3, 1, 2).float(), 2)
        p3 = ops.maxpool(x, 3, 2, 1)
        ref_p3 = F.max_pool2d(x.permute(0, 3, 1, 2).float(), 3, 2, 1)
        up = ops.upsample2(x)
        ref_up = F.interpolate(x.permute(0, 3, 1, 2).float(), scale_factor=2)
        nchw = ops.nhwc_to_nchw_f32(x, 7)
        back = ops.nchw_to_nhwc_f16(ref_up.contiguous(), 64)
        errs = [
            (p2.permute(0, 3, 1, 2).float() - ref_p2).abs().max().item(),
            (p3.permute(0, 3, 1, 2).float() - ref_p3).abs().max().item(),
            (up.permute(0, 3, 1, 2).float() - ref_up).abs().max().item(),
            (nchw - x[..., :7].permute(0, 3, 1, 2).float()).abs().max().item(),
            (back.float() - up.float()).abs().max().item(),
        ]
        torch.cuda.synchronize()
        res.update(ok=max(errs) == 0.0, max_abs=max(errs), detail=errs)
        return res
    else:
        x = (torch.randn((B, H, W, Cin), device=dev, generator=g) * 0.5).half()
        ksz = 1 if mode.startswith("1x1") else 3
        stride = 2 if mode.endswith("s2") else 1
        w = torch.randn((Cout, Cin, ksz, ksz), device=dev, generator=g) * (1.0 / (Cin * ksz * ksz) ** 0.5)
        b = torch.randn((Cout,), device=dev, generator=g) * 0.1
        pad = ksz // 2
        Ho = (H + 2 * pad - ksz) // stride + 1
        Wo = (W + 2 * pad - ksz) // stride + 1
        rs = [(r, s) for r in range(ksz) for s in range(ksz)]
        taps = [(r - pad, s - pad) for r, s in rs]
        xr = x.permute(0, 3, 1, 2).float()
        ref = F.conv2d(xr, w.half().float(), b, stride=stride, padding=pad)
        if mode == "head":
            wp = ops.pack_conv_weight(w, rs, cout_pad=16)
            got = ops.conv_taps(x, wp, ops.pad_bias(b, 16, dev), taps, Ho, Wo, head_cout=Cout)
        else:
            wp = ops.pack_conv_weight(w, rs)
            residual = None
            relu = mode in ("3x3_res",)
            if mode == "3x3_res":
                residual = (torch.randn((B, Ho, Wo, wp.shape[1]), device=dev, generator=g)).half()
                ref = F.relu(ref + residual[..., :Cout].permute(0, 3, 1, 2).float())
            y = ops.conv_taps(x, wp, ops.pad_bias(b, wp.shape[1], dev), taps, Ho, Wo, stride=stride,
                              relu=relu, residual=residual)
            got = y[..., :Cout].permute(0, 3, 1, 2).float()
    torch.cuda.synchronize()
    diff = (got - ref).abs()
    scale = ref.abs().max().item()
    res.update(max_abs=diff.max().item(), ref_max=scale, mean_abs=diff.mean().item(),
               finite=bool(torch.isfinite(got).all().item()))
    # fp16 output rounding: tolerance relative to magnitude
    tol = 2e-3 * max(scale, 1.0) if mode != "head" else 1e-4 * max(scale, 1.0)
    res["ok"] = bool(res["finite"] and res["max_abs"] <= tol)
    if not res["ok"]:
        bad = (diff > tol).nonzero()
        res["n_bad"] = int(bad.shape[0])
        res["first_bad"] = bad[:8].tolist()
        res["got_sample"] = got.flatten()[:8].tolist()
        res["ref_sample"] = ref.flatten()[:8].tolist()
    # timing
    if name.startswith("big_"):
        for _ in range(3):
            ops.conv_taps(x, wp, ops.pad_bias(b, wp.shape[1], dev), taps, Ho, Wo, relu=True)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        bb = ops.pad_bias(b, wp.shape[1], dev)
        e0.record()
        n = 10
        for _ in range(n):
            ops.conv_taps(x, wp, bb, taps, Ho, Wo, relu=True)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        flops = 2.0 * B * Ho * Wo * Cout * Cin * ksz * ksz
        res["ms"] = ms
        res["tflops"] = flops / ms / 1e9
    return res


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--case":
        try:
            r = run_case(sys.argv[2])
        except Exception as e:  # noqa
            r = {"name": sys.argv[2], "ok": False, "error": repr(e)[:500]}
        print("RESULT " + json.dumps(r))
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    only = sys.argv[1:] if len(sys.argv) > 1 else None
    results = []
    for c in CASES:
        if only and c[0] not in only:
            continue
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", c[0]],
                               capture_output=True, text=True, timeout=45)
            line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            r = json.loads(line[-1][7:]) if line else {"name": c[0], "ok": False, "error": "no result",
                                                        "stderr": p.stderr[-800:], "stdout": p.stdout[-800:]}
        except subprocess.TimeoutExpired:
            r = {"name": c[0], "ok": False, "error": "timeout"}
        r["wall_s"] = round(time.time() - t0, 1)
        print(json.dumps(r), flush=True)
        results.append(r)
        with open(os.path.join(ROOT, "gpurun_out", "conv_check.json"), "w") as fh:
            json.dump(results, fh, indent=1)
    print("PASS %d / %d" % (sum(1 for r in results if r.get("ok")), len(results)))


if __name__ == "__main__":
    main()

"""Per-layer A/B timing of the vgg-Q conv layers (B=128) for one build of libdreamb200.so.
    DREAMB200_LIB=path/to/variant.so python tools/layer_bench.py [out.json] [case ...]
Every case is launched `iters` times back to back (CUDA events around each launch, median reported); a 256 MB
buffer is rewritten between cases so no case starts with its input in L2."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dream_b200 import ops

CASES = {                      # name: (H, W, Cin, Cout, pool, head)
    "c64_64_400_pool": (400, 400, 64, 64, "only", False),
    "c64_128_200": (200, 200, 64, 128, None, False),
    "c128_128_200_pool": (200, 200, 128, 128, "only", False),
    "c128_256_100": (100, 100, 128, 256, None, False),
    "c256_256_100": (100, 100, 256, 256, None, False),
    "c256_256_100_pool": (100, 100, 256, 256, "only", False),
    "c256_512_50": (50, 50, 256, 512, None, False),
    "c512_512_50": (50, 50, 512, 512, None, False),
    "c512_512_25": (25, 25, 512, 512, None, False),
    "c128_64_100": (100, 100, 128, 64, None, False),
    "c64_64_100": (100, 100, 64, 64, None, False),
    "head_64_7_100": (100, 100, 64, 7, None, True),
}


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else None
    names = sys.argv[2:] or list(CASES)
    B, iters = 128, 12
    g = torch.Generator(device="cuda").manual_seed(0)
    flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
    res = {}
    for name in names:
        H, W, Cin, Cout, pool, head = CASES[name]
        x = (torch.randn((B, H, W, Cin), device="cuda", generator=g) * 0.5).half()
        w = torch.randn((Cout, Cin, 3, 3), device="cuda", generator=g) * (1.0 / (Cin * 9) ** 0.5)
        rs = [(r, s) for r in range(3) for s in range(3)]
        if head:
            wp, bp = ops.pack_conv_weight(w, rs, cout_pad=16), ops.pad_bias(None, 16, "cuda")
            run = lambda: ops.conv_taps(x, wp, bp, ops.TAPS_3x3, H, W, head_cout=Cout)
        else:
            wp, bp = ops.pack_conv_weight(w, rs), ops.pad_bias(None, ops.round_up(Cout, 64), "cuda")
            run = lambda: ops.conv_taps(x, wp, bp, ops.TAPS_3x3, H, W, relu=True, pool=pool)
        for _ in range(3):
            run()
        flush.fill_(1.0)
        evs = []
        for _ in range(iters):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in evs)
        ms = ts[len(ts) // 2]
        res[name] = {"ms": ms, "min": ts[0], "tflops": 2.0 * B * H * W * Cout * Cin * 9 / ms / 1e9}
        del x
    print(json.dumps({"lib": os.environ.get("DREAMB200_LIB", "default"), "cases": res}))
    if out:
        json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()

"""Times dreamb200_first_conv3x3 / _u8 at the bench shape (B=128, 400x400); DREAMB200_FC_STAGED=0 selects the
global-gather producers.  python tools/gpu_first_conv_bench.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dream_b200 import ops

B, H, W = 128, 400, 400
g = torch.Generator(device="cuda").manual_seed(0)
frames = torch.randint(0, 256, (2, B, H, W, 3), device="cuda", generator=g, dtype=torch.uint8)
x = torch.rand((2, B, 3, H, W), device="cuda", generator=g) * 2 - 1
w = torch.randn((64, 3, 3, 3), device="cuda", generator=g) * 0.2
b = torch.randn((64,), device="cuda", generator=g) * 0.1
wp, bp = ops.pack_first_weight(w, 64), ops.pad_bias(b, 64, "cuda")
norm = ((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))


def timed(fn, iters=20):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i & 1)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


out = {"staged": os.environ.get("DREAMB200_FC_STAGED", "1")}
ms = timed(lambda i: ops.first_conv3x3(x[i], wp, bp))
out["fp32"] = {"ms": ms, "GBps": (B * H * W * (12 + 128)) / ms / 1e6}
ms = timed(lambda i: ops.first_conv3x3(frames[i], wp, bp, u8_norm=norm))
out["u8"] = {"ms": ms, "GBps": (B * H * W * (3 + 128)) / ms / 1e6}
print(json.dumps(out))

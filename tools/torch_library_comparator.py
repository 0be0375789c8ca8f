"""Secondary comparator of SURVEY.md 8(d): the reference's library path (torch ops -> cuDNN) on the
SAME B200, same layer shapes and batch as bench.py's vgg-Q workloads.  It is not the contract (the
contract is the CPU fp32 reference, BASELINE.json config 1) and not part of bench.py's JSON line;
it answers "what would the unmodified reference get from its stock kernels on this GPU".

The net below is a plain nn.Sequential with the vgg-Q layer shapes (dream/models.py:557-827);
weights are random -- only the time is of interest.  Three settings:
  fp32_tf32    : torch defaults (cuDNN may use TF32), NCHW           -- what the reference runs
  fp16_nhwc    : channels_last + autocast fp16                        -- the tuned library path
  fp16_nhwc train: forward + MSE + backward (no optimizer), same setting

  python tools/torch_library_comparator.py [batch] > profiles/r01_torch_library_comparator.json
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn

from dream_b200.models import VGG_TRUNK


def vgg_q(n_kp=7):
    layers, cin = [], 3
    for bi, (_, idxs, ch) in enumerate(VGG_TRUNK):
        if bi:
            layers.append(nn.MaxPool2d(2))
        for _j in idxs:
            layers += [nn.Conv2d(cin, ch, 3, padding=1), nn.ReLU(inplace=True)]
            cin = ch
    for ci, mid, co in ((512, 256, 256), (256, 128, 64)):
        layers += [nn.Upsample(scale_factor=2), nn.Conv2d(ci, mid, 3, padding=1), nn.ReLU(inplace=True),
                   nn.Conv2d(mid, co, 3, padding=1)]
    layers += [nn.Conv2d(64, 64, 3, padding=1), nn.ReLU(inplace=True), nn.Conv2d(64, 32, 3, padding=1),
               nn.ReLU(inplace=True), nn.Conv2d(32, n_kp, 3, padding=1)]
    return nn.Sequential(*layers)


def timed(fn, warm=3, iters=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    torch.manual_seed(0)
    torch.backends.cudnn.benchmark = True
    net = vgg_q().cuda()
    x = torch.rand((B, 3, 400, 400), device="cuda") * 2 - 1
    out = {"batch": B, "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(),
           "gflop_per_img_fwd": 141.7824}

    def rec(name, fn, flop_mult=1.0):
        try:
            ms = timed(fn)
            out[name] = {"ms": round(ms, 3), "img_per_s": round(B / ms * 1e3, 1),
                         "tflops": round(B * 141.7824e9 * flop_mult / ms / 1e9, 1)}
        except RuntimeError as e:                      # e.g. CUDA OOM at this batch: report, don't die
            out[name] = {"error": str(e).splitlines()[0][:200]}
            torch.cuda.empty_cache()

    net.eval()
    with torch.no_grad():
        rec("infer_fp32_tf32_nchw", lambda: net(x))
        net_cl = net.to(memory_format=torch.channels_last)
        x_cl = x.contiguous(memory_format=torch.channels_last)

        def f16():
            with torch.autocast("cuda", dtype=torch.float16):
                return net_cl(x_cl)
        rec("infer_fp16_nhwc", f16)

    net_cl.train()
    tgt = torch.zeros((B, 7, 100, 100), device="cuda")

    def train_step():
        for p in net_cl.parameters():
            p.grad = None
        with torch.autocast("cuda", dtype=torch.float16):
            y = net_cl(x_cl)
        loss = torch.nn.functional.mse_loss(y.float(), tgt)
        loss.backward()
    rec("train_fp16_nhwc_fwd_bwd", train_step, flop_mult=424.79 / 141.7824)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

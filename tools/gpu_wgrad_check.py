"""Bring-up harness for the wgrad tensor-core kernel and the streaming backward kernels; one
subprocess per case (a faulting kernel kills only that case).  Writes gpurun_out/wgrad_check.json."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = []
for (B, H, W, Ci, Co) in [(1, 8, 16, 64, 64), (2, 16, 24, 64, 64), (2, 16, 24, 64, 128), (3, 25, 25, 128, 256),
                          (2, 50, 37, 256, 128), (2, 13, 13, 512, 512)]:
    CASES.append(("wgrad_%dx%dx%d_%d_%d" % (B, H, W, Ci, Co), 0, B, H, W, Ci, Co))
CASES.append(("stream", 0, 0, 0, 0, 0, 0))
CASES.append(("wgrad_auto_big_32x100x100_256_256", 0, 32, 100, 100, 256, 256))
CASES.append(("wgrad_auto_big_8x400x400_64_64", 0, 8, 400, 400, 64, 64))


def run_case(name):
    import torch
    import torch.nn.functional as F
    case = [c for c in CASES if c[0] == name][0]
    _, bx, B, H, W, Ci, Co = case
    if bx:
        os.environ["DREAMB200_WGRAD_BX"] = str(bx)
    from dream_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(3)
    res = {"name": name}
    if name == "stream":
        x = torch.randn((2, 26, 31, 64), device="cuda", generator=g).half()
        xr = x.permute(0, 3, 1, 2).float().requires_grad_(True)
        yp = F.max_pool2d(xr, 2)
        dy = torch.randn(yp.shape, device="cuda", generator=g).half()
        yp.backward(dy.float())
        dx = ops.maxpool2_bwd(x, dy.permute(0, 2, 3, 1).contiguous())
        e1 = (dx.permute(0, 3, 1, 2).float() - xr.grad).abs().max().item()
        xr.grad = None
        yu = F.interpolate(xr, scale_factor=2)
        dyu = torch.randn(yu.shape, device="cuda", generator=g).half()
        yu.backward(dyu.float())
        dxu = ops.upsample2_bwd(dyu.permute(0, 2, 3, 1).contiguous())
        e2 = (dxu.permute(0, 3, 1, 2).float() - xr.grad).abs().max().item()
        y = torch.relu(x)
        d = torch.randn(x.shape, device="cuda", generator=g).half()
        ref_mask = d.float() * (y.float() > 0)
        got = ops.relu_mask_(d.clone(), y)
        e3 = (got.float() - ref_mask).abs().max().item()
        sc = torch.full((1,), 4.0, device="cuda")
        e3 += (ops.scale_mask_(d.clone(), y, sc).float() - 4 * ref_mask).abs().max().item()
        e3 += abs(ops.absmax(d).item() - d.float().abs().max().item())
        db = ops.bias_grad(got)
        e4 = (db - ref_mask.sum(dim=(0, 1, 2))).abs().max().item()
        e5 = 0.0
        torch.cuda.synchronize()
        res.update(ok=(e1 == 0 and e2 < 2e-2 and e3 == 0 and e4 < 1e-2 and e5 == 0), detail=[e1, e2, e3, e4, e5])
        return res
    x = (torch.randn((B, H, W, Ci), device="cuda", generator=g) * 0.5).half()
    dy = (torch.randn((B, H, W, Co), device="cuda", generator=g) * 0.5).half()
    dw = ops.wgrad(dy, x, ops.TAPS_3x3)
    torch.cuda.synchronize()
    # reference: explicit shifted einsum in fp32 (no cuDNN)
    xp = F.pad(x.float(), (0, 0, 1, 1, 1, 1))
    ref = torch.empty((9, Co, Ci), device="cuda")
    for t, (dyy, dxx) in enumerate(ops.TAPS_3x3):
        xs = xp[:, 1 + dyy:1 + dyy + H, 1 + dxx:1 + dxx + W, :]
        ref[t] = torch.einsum("bhwo,bhwi->oi", dy.float(), xs)
    err = (dw - ref).abs().max().item()
    scale = ref.abs().max().item()
    res.update(max_abs=err, ref_max=scale, ok=bool(err <= 2e-3 * scale))
    if "big" in name:
        for _ in range(2):
            ops.wgrad(dy, x, ops.TAPS_3x3)
        ops.PROFILE = []
        for _ in range(5):
            ops.wgrad(dy, x, ops.TAPS_3x3)
        torch.cuda.synchronize()
        ms = sum(e0.elapsed_time(e1) for _, _, e0, e1 in ops.PROFILE) / 5
        res["ms"] = ms
        res["tflops"] = 2.0 * B * H * W * Co * Ci * 9 / ms / 1e9
    return res


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--case":
        try:
            r = run_case(sys.argv[2])
        except Exception as e:  # noqa
            r = {"name": sys.argv[2], "ok": False, "error": repr(e)[:300]}
        print("RESULT " + json.dumps(r))
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    results = []
    for c in CASES:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", c[0]],
                               capture_output=True, text=True, timeout=45)
            line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            r = json.loads(line[-1][7:]) if line else {"name": c[0], "ok": False, "error": "no result",
                                                        "stderr": p.stderr[-400:], "stdout": p.stdout[-400:]}
        except subprocess.TimeoutExpired:
            r = {"name": c[0], "ok": False, "error": "timeout"}
        r["wall_s"] = round(time.time() - t0, 1)
        print(json.dumps(r), flush=True)
        results.append(r)
        json.dump(results, open(os.path.join(ROOT, "gpurun_out", "wgrad_check.json"), "w"), indent=1)
    print("PASS %d / %d" % (sum(1 for r in results if r.get("ok")), len(results)))


if __name__ == "__main__":
    main()

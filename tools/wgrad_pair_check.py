"""Bring-up harness for the CTA-pair weight-gradient kernel (wgrad_pair.cu; default since round 2, DREAMB200_WGRAD3_2SM=0 switches it off): every
(case, mode) in its own subprocess; dW is compared with the single-CTA kernel (fp32 atomics: not bit-exact, gate 1e-5
of the largest entry) and with torch autograd, and timed.   (round 2: ALL OK on a B200, profiles/r02_ab_pair_kernels.txt)"""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CASES = {  # name: (B, H, W, Cin, Cout, timed)
    "s128_256_25": (3, 25, 25, 128, 256, False),
    "s256_256_37x50": (2, 50, 37, 256, 256, False),
    "s512_512_13": (2, 13, 13, 512, 512, False),
    "s256_512_1x9": (2, 1, 9, 256, 512, False),
    "b256_256_100": (128, 100, 100, 256, 256, True),
    "b512_512_50": (128, 50, 50, 512, 512, True),
    "b512_512_25": (128, 25, 25, 512, 512, True),
    "b128_256_100": (128, 100, 100, 128, 256, True),
}


def run_case(name, out_path):
    import torch
    from dream_b200 import ops
    B, H, W, Ci, Co, timed = CASES[name]
    g = torch.Generator(device="cuda").manual_seed(3)
    x = (torch.randn((B, H, W, Ci), device="cuda", generator=g) * 0.5).half()
    dy = (torch.randn((B, H, W, Co), device="cuda", generator=g) * 0.5).half()
    dw = ops.wgrad(dy, x, ops.TAPS_3x3)                         # [9, Co, Ci]
    torch.cuda.synchronize()
    res = {"name": name, "finite": bool(torch.isfinite(dw).all())}
    if not timed:
        xr = x.permute(0, 3, 1, 2).float().requires_grad_(True)
        w = torch.zeros((Co, Ci, 3, 3), device="cuda", requires_grad=True)
        torch.backends.cudnn.allow_tf32 = False
        y = torch.nn.functional.conv2d(xr, w, padding=1)
        y.backward(dy.permute(0, 3, 1, 2).float())
        ref = w.grad.permute(2, 3, 0, 1).reshape(9, Co, Ci)
        res["rel_vs_torch"] = float((dw - ref).abs().max() / ref.abs().max())
    else:
        for _ in range(2):
            ops.wgrad(dy, x, ops.TAPS_3x3)
        evs = []
        for _ in range(6):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); ops.wgrad(dy, x, ops.TAPS_3x3); e1.record(); evs.append((e0, e1))
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in evs)
        res["ms"] = ts[len(ts) // 2]
        res["tflops"] = 2.0 * B * H * W * Co * Ci * 9 / res["ms"] / 1e9
    torch.save(dw.cpu(), out_path)
    print("RESULT " + json.dumps(res))


def main():
    if len(sys.argv) > 3 and sys.argv[1] == "--case":
        return run_case(sys.argv[2], sys.argv[3])
    import torch
    ok_all = True
    for name in (sys.argv[1:] or list(CASES)):
        got = {}
        for mode in ("0", "1"):
            env = dict(os.environ, DREAMB200_WGRAD3_2SM=mode)
            outp = "/tmp/wgp_%s_%s.pt" % (name, mode)
            t0 = time.time()
            try:
                p = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", name, outp], env=env,
                                   capture_output=True, text=True, timeout=240)
                line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
                got[mode] = json.loads(line[-1][7:]) if line else {"error": (p.stderr[-600:] + p.stdout[-300:])}
            except subprocess.TimeoutExpired:
                got[mode] = {"error": "timeout"}
            got[mode]["wall"] = round(time.time() - t0, 1)
        rec = {"name": name, "ref": got["0"], "pair": got["1"]}
        if "error" not in got["0"] and "error" not in got["1"]:
            a = torch.load("/tmp/wgp_%s_0.pt" % name); b = torch.load("/tmp/wgp_%s_1.pt" % name)
            rec["rel_diff"] = float((a - b).abs().max() / a.abs().max())
            rec["ok"] = rec["rel_diff"] <= (5e-5 if CASES[name][5] else 1e-5)   # B = 128: ~1e6-term fp32 sums, different k-blocking
        ok_all = ok_all and rec.get("ok", False)
        print(json.dumps(rec), flush=True)
    print("ALL OK" if ok_all else "MISMATCH / ERROR")


if __name__ == "__main__":
    main()

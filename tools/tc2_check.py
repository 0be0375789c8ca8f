"""Bring-up harness for the CTA-pair wide-layer kernel (conv_tc2.cu; default since round 2, DREAMB200_TC2=0 switches it off): every (case, mode) runs
in its own subprocess (a trapped kernel kills only that run); outputs are compared bit for bit with conv_tc and the
big cases are timed.  python tools/tc2_check.py [case ...]      (round 2: ALL IDENTICAL on a B200, profiles/r02_ab_pair_kernels.txt)"""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {  # name: (B, H, W, Cin, Cout, kind, timed)   kind: 3x3 | pool | s2 | 1x1res | odd
    "s256_256_50": (2, 50, 50, 256, 256, "3x3", False),
    "s128_256_100_pool": (2, 100, 100, 128, 256, "pool", False),
    "s256_512_26_s2": (2, 26, 30, 256, 512, "s2", False),
    "s256_1024_25_1x1res": (2, 25, 25, 256, 1024, "1x1res", False),
    "s512_512_25_odd": (1, 25, 25, 512, 512, "3x3", False),          # 5 M-tiles: the last pair has one real tile
    "s64_256_9x7": (1, 7, 9, 64, 256, "3x3", False),                 # a single M-tile: falls back to conv_tc
    "b256_256_100": (128, 100, 100, 256, 256, "3x3", True),
    "b512_512_50": (128, 50, 50, 512, 512, "3x3", True),
    "b512_512_25": (128, 25, 25, 512, 512, "3x3", True),
    "b256_256_100_pool": (128, 100, 100, 256, 256, "pool", True),
}


def run_case(name, out_path):
    import torch
    from dream_b200 import ops
    B, H, W, Cin, Cout, kind, timed = CASES[name]
    g = torch.Generator(device="cuda").manual_seed(11)
    x = (torch.randn((B, H, W, Cin), device="cuda", generator=g) * 0.5).half()
    ksz = 1 if kind == "1x1res" else 3
    stride = 2 if kind == "s2" else 1
    pad = ksz // 2
    w = torch.randn((Cout, Cin, ksz, ksz), device="cuda", generator=g) * (1.0 / (Cin * ksz * ksz) ** 0.5)
    bias = torch.randn((Cout,), device="cuda", generator=g) * 0.1
    rs = [(r, s) for r in range(ksz) for s in range(ksz)]
    taps = [(r - pad, s - pad) for r, s in rs]
    Ho = (H + 2 * pad - ksz) // stride + 1
    Wo = (W + 2 * pad - ksz) // stride + 1
    wp, bp = ops.pack_conv_weight(w, rs), ops.pad_bias(bias, Cout, "cuda")
    kw = {"relu": True, "stride": stride}
    if kind == "pool":
        kw["pool"] = "both"
    if kind == "1x1res":
        kw["residual_f32"] = torch.randn((B, Ho, Wo, Cout), device="cuda", generator=g)
        kw["y_f32"] = torch.empty((B, Ho, Wo, Cout), device="cuda")
    run = lambda: ops.conv_taps(x, wp, bp, taps, Ho, Wo, **kw)
    y = run()
    torch.cuda.synchronize()
    outs = [t for t in (y if isinstance(y, tuple) else (y,)) if t is not None]
    if kind == "1x1res":
        outs.append(kw["y_f32"])
    res = {"name": name, "finite": all(bool(torch.isfinite(t.float()).all()) for t in outs)}
    if timed:
        for _ in range(2):
            run()
        evs = []
        for _ in range(8):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); evs.append((e0, e1))
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in evs)
        res["ms"] = ts[len(ts) // 2]
        res["tflops"] = 2.0 * B * Ho * Wo * Cout * Cin * ksz * ksz / res["ms"] / 1e9
        res["sum"] = [float(t.float().sum()) for t in outs]
        outs = [t[:2].contiguous() for t in outs]
    torch.save([t.cpu() for t in outs], out_path)
    print("RESULT " + json.dumps(res))


def main():
    if len(sys.argv) > 3 and sys.argv[1] == "--case":
        return run_case(sys.argv[2], sys.argv[3])
    import torch
    names = sys.argv[1:] or list(CASES)
    ok_all = True
    for name in names:
        got = {}
        for mode in ("0", "1"):
            env = dict(os.environ, DREAMB200_TC2=mode)
            outp = "/tmp/tc2_%s_%s.pt" % (name, mode)
            t0 = time.time()
            try:
                p = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", name, outp], env=env,
                                   capture_output=True, text=True, timeout=240)
                line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
                got[mode] = json.loads(line[-1][7:]) if line else {"error": (p.stderr[-600:] + p.stdout[-300:])}
            except subprocess.TimeoutExpired:
                got[mode] = {"error": "timeout"}
            got[mode]["wall"] = round(time.time() - t0, 1)
        rec = {"name": name, "ref": got["0"], "tc2": got["1"]}
        if "error" not in got["0"] and "error" not in got["1"]:
            a = torch.load("/tmp/tc2_%s_0.pt" % name); b = torch.load("/tmp/tc2_%s_1.pt" % name)
            rec["max_abs_diff"] = max(float((u.float() - v.float()).abs().max()) for u, v in zip(a, b))
            rec["identical"] = all(torch.equal(u, v) for u, v in zip(a, b))
            if "sum" in got["0"]:
                rec["sum_match"] = got["0"]["sum"] == got["1"]["sum"]
        ok_all = ok_all and rec.get("identical", False)
        print(json.dumps(rec), flush=True)
    print("ALL IDENTICAL" if ok_all else "MISMATCH / ERROR")


if __name__ == "__main__":
    main()

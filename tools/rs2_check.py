"""Bring-up harness for the CTA-pair slab kernel (conv_rs2.cu): every (case, DREAMB200_RS2 mode) runs in its own
subprocess (a trapped kernel kills only that run); outputs are compared bit for bit with the single-CTA path and
timed.  python tools/rs2_check.py [case ...]"""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {  # name: (B, H, W, Cin, Cout, pool, timed)
    "s64_64_32": (2, 32, 32, 64, 64, None, False),
    "s64_64_40x56_pool": (2, 40, 56, 64, 64, "only", False),
    "s64_64_100_both": (2, 100, 100, 64, 64, "both", False),
    "s64_128_64": (2, 64, 40, 64, 128, None, False),
    "s128_128_48_pool": (2, 48, 48, 128, 128, "only", False),
    "s128_64_100": (2, 100, 100, 128, 64, None, False),
    "s256_128_50": (2, 50, 50, 256, 128, None, False),
    "g128_128_64_bwd": (2, 64, 48, 128, 128, "bwd", False),     # backward-pass options: gate, out_scale, colsum, absmax
    "g64_64_48_bwd": (3, 48, 40, 64, 64, "bwd", False),
    "b64_64_400_pool": (128, 400, 400, 64, 64, "only", True),
    "b64_128_200": (128, 200, 200, 64, 128, None, True),
    "b128_128_200_pool": (128, 200, 200, 128, 128, "only", True),
    "b128_64_100": (128, 100, 100, 128, 64, None, True),
    "b64_64_100": (128, 100, 100, 64, 64, None, True),
}


def run_case(name, out_path):
    import torch
    from dream_b200 import ops
    B, H, W, Cin, Cout, pool, timed = CASES[name]
    g = torch.Generator(device="cuda").manual_seed(7)
    x = (torch.randn((B, H, W, Cin), device="cuda", generator=g) * 0.5).half()
    w = torch.randn((Cout, Cin, 3, 3), device="cuda", generator=g) * (1.0 / (Cin * 9) ** 0.5)
    bias = torch.randn((Cout,), device="cuda", generator=g) * 0.1
    rs = [(r, s) for r in range(3) for s in range(3)]
    wp, bp = ops.pack_conv_weight(w, rs), ops.pad_bias(bias, ops.round_up(Cout, 64), "cuda")
    if pool == "bwd":
        Cp = ops.round_up(Cout, 64)
        gate = torch.randn((B, H, W, Cp), device="cuda", generator=g).half()
        scale = torch.full((1,), 0.5, device="cuda")
        colsum = torch.zeros((Cp,), device="cuda"); amax = torch.zeros((1,), device="cuda")
        y = ops.conv_taps(x, wp, None, ops.TAPS_3x3, H, W, gate=gate, out_scale=scale, colsum=colsum, absmax=amax)
        torch.cuda.synchronize()
        # column sums are float atomics (order differs run to run): compare them rounded to 3 significant digits
        cs = torch.round(colsum / colsum.abs().max() * 1000.0)
        outs = [y, cs, amax]
        torch.save([t.cpu() for t in outs], out_path)
        print("RESULT " + json.dumps({"name": name, "finite": bool(torch.isfinite(y.float()).all())}))
        return
    run = lambda: ops.conv_taps(x, wp, bp, ops.TAPS_3x3, H, W, relu=True, pool=pool)
    y = run()
    torch.cuda.synchronize()
    outs = [t for t in (y if isinstance(y, tuple) else (y,)) if t is not None]
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
        res["sum"] = [float(t.float().sum()) for t in outs]
        res["absmax"] = [float(t.float().abs().max()) for t in outs]
        # keep a slice for the bit comparison (the full tensors are GBs)
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
        for mode in ("0", "15"):
            env = dict(os.environ, DREAMB200_RS2=mode)
            outp = "/tmp/rs2_%s_%s.pt" % (name, mode)
            t0 = time.time()
            try:
                p = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", name, outp], env=env,
                                   capture_output=True, text=True, timeout=240)
                line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
                got[mode] = json.loads(line[-1][7:]) if line else {"error": (p.stderr[-600:] + p.stdout[-300:])}
            except subprocess.TimeoutExpired:
                got[mode] = {"error": "timeout"}
            got[mode]["wall"] = round(time.time() - t0, 1)
        rec = {"name": name, "ref": got["0"], "rs2": got["15"]}
        if "error" not in got["0"] and "error" not in got["15"]:
            a = torch.load("/tmp/rs2_%s_0.pt" % name); b = torch.load("/tmp/rs2_%s_15.pt" % name)
            rec["max_abs_diff"] = max(float((u.float() - v.float()).abs().max()) for u, v in zip(a, b))
            rec["identical"] = all(torch.equal(u, v) for u, v in zip(a, b))
            if "sum" in got["0"]:
                rec["sum_match"] = got["0"]["sum"] == got["15"]["sum"]
        ok_all = ok_all and rec.get("identical", False)
        print(json.dumps(rec), flush=True)
    print("ALL IDENTICAL" if ok_all else "MISMATCH / ERROR")


if __name__ == "__main__":
    main()

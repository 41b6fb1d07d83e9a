#!/usr/bin/env python
"""Fused search kernel (awqk_scale_search) against the stand-alone stages (awqk_fakequant_delta -> HBM ->
awqk_sqerr_gemm): scores must agree to fp64-atomic ordering, and the two are timed side by side.
    python tools/probe_fused.py [--shapes 4096x4096,14336x4096,4096x14336] [--tokens 2048] [--reps 3]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "awq-converter_b200"))
import torch
from awq_quantizer import _native as N
from awq_quantizer.quantization import search as S

ap = argparse.ArgumentParser()
ap.add_argument("--shapes", default="256x512,512x1024,1024x1024,4096x4096,14336x4096,4096x14336")
ap.add_argument("--tokens", type=int, default=2048)
ap.add_argument("--n_grid", type=int, default=20)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--ring", default="pref", help="comma list of: pref | min | <bytes>")
ap.add_argument("--sustain-ms", type=float, default=0.0, help="time each variant over at least this long (power-capped regime)")
ap.add_argument("--no-staged", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda:0")
T, n = args.tokens, args.n_grid
gen = torch.Generator(device=dev).manual_seed(1)


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for spec in args.shapes.split(","):
    C, K = (int(v) for v in spec.split("x"))
    w = (torch.randn((C, K), generator=gen, device=dev) * 0.02).to(torch.bfloat16)
    gain = torch.exp(torch.randn(K, generator=gen, device=dev))
    x = (torch.randn((T, K), generator=gen, device=dev) * gain).to(torch.bfloat16)
    st = N.stream_ptr(dev)
    _, grid, xb = S.activation_grid(x, n, st)
    pref, mn = S.workspace_bytes(C, K, T, n)
    out = {}

    def staged():
        out["s"] = S.search_device_staged(w, x, grid, bits=4, group_size=128, symmetric=False)

    flops = 2.0 * T * C * K * n
    s_ms = None
    if not args.no_staged:
        s_ms = timed(staged, args.reps)
    for ring in args.ring.split(","):
        if ring == "pref":
            nbytes = pref
        elif ring == "min":
            nbytes = mn
        else:                       # bytes
            nbytes = max(mn, int(ring))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)

        def fused():
            out["r"] = S.scale_search(w, xb, grid, bits=4, group_size=128, symmetric=False, workspace=ws)

        f_ms = timed(fused, args.reps)
        rec = {"shape": [C, K], "tokens": T, "n_grid": n, "ring": ring, "workspace_mb": round(nbytes / 2**20, 1),
               "fused_ms": round(f_ms, 3), "fused_tflops": round(flops / f_ms / 1e9, 1)}
        if args.sustain_ms:
            reps = max(3, int(args.sustain_ms / f_ms))
            f_sus = timed(fused, reps)
            rec.update({"fused_sustained_ms": round(f_sus, 3), "fused_sustained_tflops": round(flops / f_sus / 1e9, 1),
                        "sustained_reps": reps})
        if s_ms is not None:
            ef = (out["r"]["err_mean"] * float(T * C)).cpu()
            es = out["s"].cpu()
            rec.update({"staged_ms": round(s_ms, 3), "staged_tflops": round(flops / s_ms / 1e9, 1),
                        "max_rel_diff_scores": float(((ef - es).abs() / es).max()),
                        "best_fused": int(out["r"]["best_idx"]), "best_staged": int(torch.argmin(es))})
        print(json.dumps(rec), flush=True)
        del ws

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
ap.add_argument("--ring", default="pref", help="pref | min | <bytes>")
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
    nbytes = pref if args.ring == "pref" else mn if args.ring == "min" else int(args.ring)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out = {}

    def fused():
        out["r"] = S.scale_search(w, xb, grid, bits=4, group_size=128, symmetric=False, workspace=ws)

    def staged():
        out["s"] = S.search_device_staged(w, x, grid, bits=4, group_size=128, symmetric=False)

    f_ms = timed(fused, args.reps)
    s_ms = timed(staged, args.reps)
    ef = (out["r"]["err_mean"] * float(T * C)).cpu()
    es = out["s"].cpu()
    rel = float(((ef - es).abs() / es).max())
    flops = 2.0 * T * C * K * n
    print(json.dumps({"shape": [C, K], "tokens": T, "n_grid": n, "workspace_mb": round(nbytes / 2**20, 1),
                      "fused_ms": round(f_ms, 3), "staged_ms": round(s_ms, 3),
                      "fused_tflops": round(flops / f_ms / 1e9, 1), "staged_tflops": round(flops / s_ms / 1e9, 1),
                      "max_rel_diff_scores": rel, "best_fused": int(out["r"]["best_idx"]),
                      "best_staged": int(torch.argmin(es))}), flush=True)

#!/usr/bin/env python
"""Sustained (power-capped) throughput and clocks of: the stand-alone tcgen05 GEMM (delta precomputed), the fused
search kernel, and torch.matmul (cuBLAS) on the same problem; NVML clocks / power sampled while each runs.
    python tools/probe_sustained.py [--shapes 4096x4096,4096x14336] [--seconds 2]"""
import argparse, json, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "awq-converter_b200"))
import torch
import pynvml
from awq_quantizer import _native as N
from awq_quantizer.quantization import search as S

ap = argparse.ArgumentParser()
ap.add_argument("--shapes", default="4096x4096,4096x14336")
ap.add_argument("--tokens", type=int, default=2048)
ap.add_argument("--seconds", type=float, default=2.0)
ap.add_argument("--what", default="gemm,fused,cublas")
args = ap.parse_args()
dev = torch.device("cuda:0")
L = N.lib()
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
T, n = args.tokens, 20
gen = torch.Generator(device=dev).manual_seed(1)


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.stop_flag, self.rows = False, []

    def run(self):
        while not self.stop_flag:
            self.rows.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
            time.sleep(0.05)

    def summary(self):
        rows = self.rows[len(self.rows) // 3:]                 # the settled part
        if not rows:
            return {}
        mhz = sorted(r[0] for r in rows); pw = sorted(r[1] for r in rows)
        return {"sm_mhz_median": mhz[len(mhz) // 2], "power_w_median": round(pw[len(pw) // 2], 1), "samples": len(rows)}


def sustained(fn, seconds, flops):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    one = e0.elapsed_time(e1)
    reps = max(3, int(seconds * 1e3 / one))
    s = Sampler(); s.start()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    s.stop_flag = True; s.join()
    ms = e0.elapsed_time(e1) / reps
    return {"ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1), "first_ms": round(one, 4), "reps": reps, **s.summary()}


for spec in args.shapes.split(","):
    C, K = (int(v) for v in spec.split("x"))
    w = (torch.randn((C, K), generator=gen, device=dev) * 0.02).to(torch.bfloat16)
    x = (torch.randn((T, K), generator=gen, device=dev) * torch.exp(torch.randn(K, generator=gen, device=dev))).to(torch.bfloat16)
    st = N.stream_ptr(dev)
    _, grid, xb = S.activation_grid(x, n, st)
    flops = 2.0 * T * C * K * n
    rec = {"shape": [C, K], "tokens": T}
    if "gemm" in args.what or "cublas" in args.what:
        dw = torch.empty((n, C, K), dtype=torch.bfloat16, device=dev)
        N.check(L.awqk_fakequant_delta(w.data_ptr(), N.BF16, C, K, 128, 4, 0, grid.data_ptr(), n, dw.data_ptr(), st))
        err = torch.zeros(n, dtype=torch.float64, device=dev)
    if "gemm" in args.what:
        rec["gemm_alone"] = sustained(lambda: N.check(L.awqk_sqerr_gemm(xb.data_ptr(), dw.data_ptr(), T, C, K, n, err.data_ptr(), st)), args.seconds, flops)
        time.sleep(1.0)
    if "delta" in args.what:
        rec["delta_alone"] = sustained(lambda: N.check(L.awqk_fakequant_delta(w.data_ptr(), N.BF16, C, K, 128, 4, 0, grid.data_ptr(), n, dw.data_ptr(), st)), args.seconds, flops)
        time.sleep(1.0)
    if "fused" in args.what:
        ws = torch.empty(S.workspace_bytes(C, K, T, n)[0], dtype=torch.uint8, device=dev)
        rec["fused"] = sustained(lambda: S.scale_search(w, xb, grid, bits=4, group_size=128, symmetric=False, workspace=ws), args.seconds, flops)
        time.sleep(1.0)
    if "cublas" in args.what:
        out = torch.empty((T, n * C), dtype=torch.bfloat16, device=dev)
        d2 = dw.view(n * C, K)
        rec["cublas_matmul_same_shape"] = sustained(lambda: torch.matmul(xb, d2.t(), out=out), args.seconds, flops)
        time.sleep(1.0)
    print(json.dumps(rec), flush=True)
    del w, x
    if "gemm" in args.what or "cublas" in args.what:
        del dw

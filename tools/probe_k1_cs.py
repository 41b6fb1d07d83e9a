"""Timing probe for the column-scaled (final AWQ) K1 pass (column-slab mode of the TMA kernel).  CUDA events,
2 rotating inputs > L2."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "awq-converter_b200"))
import torch
from awq_quantizer import _native as N

L = N.lib()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream(dev).cuda_stream
iters = int(os.environ.get("ITERS", "20"))
tag = "slab kernel"


def timeit(fn, iters=iters, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


res = []
for C, K in ((8192, 28672), (28672, 8192), (14336, 4096), (4096, 14336), (4096, 4096), (65536, 1024)):
    n = C * K
    bufs = [(torch.randn((C, K), device=dev, dtype=torch.float32) * 0.02).to(torch.bfloat16) for _ in range(2 if n * 2 > (200 << 20) else 8)]
    s = torch.exp(0.5 * torch.randn(K, device=dev)).float()
    for g in (128, 32):
        G = K // g
        scales = torch.empty((C, G), dtype=torch.float16, device=dev)
        zp = torch.empty((C, G), dtype=torch.int32, device=dev)
        qw = torch.empty((C, K // 8), dtype=torch.int32, device=dev)
        qz = torch.empty((C, G // 8), dtype=torch.int32, device=dev)
        for unp in (False, True):
            if unp and g != 128:
                continue
            q = torch.empty((C, K), dtype=torch.int32, device=dev) if unp else None
            def run(i):
                N.check(L.awqk_group_quant(bufs[i % len(bufs)].data_ptr(), N.BF16, C, K, g, 4, 0, N.ARITH_FP32, N.ptr(q),
                                           qw.data_ptr(), scales.data_ptr(), zp.data_ptr(), qz.data_ptr(), s.data_ptr(), st))
            t = timeit(run)
            bpe = 2 + 0.5 + 2.0 / g + 4.0 / g + 0.5 / g + (4 if unp else 0)
            res.append(dict(kernel="K1 col_scale " + tag, shape=[C, K], g=g, unpacked=unp, us=round(t * 1e6, 1),
                            gbs_bf16=round(2 * n / t / 1e9), hbm_gbs=round(bpe * n / t / 1e9)))
            print(json.dumps(res[-1]), flush=True)
            del q
    del bufs
# ---- batches: the linears of whole Llama-3-8B layers in ONE launch per 32 tensors vs one launch per tensor ----
layer = [(4096, 4096), (1024, 4096), (1024, 4096), (4096, 4096), (14336, 4096), (14336, 4096), (4096, 14336)]
for n_layers in (1, 4, 8):
    shapes = layer * n_layers
    g = 128
    items = []
    for C, K in shapes:
        w = (torch.randn((C, K), device=dev, dtype=torch.float32) * 0.02).to(torch.bfloat16)
        s = torch.exp(0.5 * torch.randn(K, device=dev)).float()
        items.append((w, C, K, s, None, torch.empty((C, K // 8), dtype=torch.int32, device=dev),
                      torch.empty((C, K // g), dtype=torch.float16, device=dev), None,
                      torch.empty((C, K // g // 8), dtype=torch.int32, device=dev)))
    n = sum(C * K for C, K in shapes)
    bpe = 2 + 0.5 + 2.0 / g + 0.5 / g
    def run_batch(i):
        N.group_quant_batch(items, N.BF16, g, 4, False, N.ARITH_FP32, st)
    def run_single(i):
        for (w, C, K, s, _, qw, sc, _, qz) in items:
            N.check(L.awqk_group_quant(w.data_ptr(), N.BF16, C, K, g, 4, 0, N.ARITH_FP32, None, qw.data_ptr(), sc.data_ptr(),
                                       None, qz.data_ptr(), s.data_ptr(), st))
    for name, fn in (("batch", run_batch), ("single", run_single)):
        t = timeit(fn, iters=10 if n_layers < 8 else 5)
        res.append(dict(kernel="K1 col_scale " + name, layers=n_layers, tensors=len(shapes), us=round(t * 1e6, 1),
                        hbm_gbs=round(bpe * n / t / 1e9), frac=round(bpe * n / t / 1e9 / 6546.6, 3)))
        print(json.dumps(res[-1]), flush=True)
    del items
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "probe_k1_cs_" + tag.split()[0] + ".json"), "w"), indent=1)

"""Quick K1/K3/K4 timing probe on the micro-bench matrix (8192 x 28672).  CUDA events, 2 rotating
input buffers (each 470 MB > L2).  Prints GB/s of bf16 weights and algorithmic HBM GB/s."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "awq-converter_b200"))
import torch
from awq_quantizer import _native as N

L = N.lib()
dev = torch.device("cuda:0")
C, K = int(os.environ.get("ROWS", "8192")), 28672
n = C * K
iters = int(os.environ.get("ITERS", "20"))
bufs = [(torch.randn((C, K), device=dev, dtype=torch.float32) * 0.02).to(torch.bfloat16) for _ in range(2)]
st = torch.cuda.current_stream(dev).cuda_stream


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
for g in (128, 64, 32):
    G = K // g
    scales = torch.empty((C, G), dtype=torch.float16, device=dev)
    zp = torch.empty((C, G), dtype=torch.int32, device=dev)
    qw = torch.empty((C, K // 8), dtype=torch.int32, device=dev)
    qz = torch.empty((C, G // 8), dtype=torch.int32, device=dev)
    for arith, aname in ((N.ARITH_NATIVE, "native"), (N.ARITH_FP32, "fp32")):
        for sym in (0, 1):
            def run(i):
                N.check(L.awqk_group_quant(bufs[i & 1].data_ptr(), N.BF16, C, K, g, 4, sym, arith, None,
                                           qw.data_ptr(), scales.data_ptr(), None, qz.data_ptr(), None, st))
            t = timeit(run)
            bpe = 2 + 0.5 + 2.0 / g + 0.5 / g
            res.append(dict(kernel="K1 pack", g=g, arith=aname, sym=sym, us=t * 1e6, gbs_bf16=2 * n / t / 1e9,
                            hbm_gbs=bpe * n / t / 1e9))
            print(res[-1], flush=True)
# reference-compatible unpacked int32 output
g = 128
G = K // g
scales = torch.empty((C, G), dtype=torch.float16, device=dev)
zp = torch.empty((C, G), dtype=torch.int32, device=dev)
q = torch.empty((C, K), dtype=torch.int32, device=dev)
def run(i):
    N.check(L.awqk_group_quant(bufs[i & 1].data_ptr(), N.BF16, C, K, g, 4, 0, N.ARITH_NATIVE, q.data_ptr(), None,
                               scales.data_ptr(), zp.data_ptr(), None, None, st))
t = timeit(run)
res.append(dict(kernel="K1 unpacked int32", g=128, us=t * 1e6, gbs_bf16=2 * n / t / 1e9, hbm_gbs=(6 + 6 / g) * n / t / 1e9))
print(res[-1], flush=True)
# K4 dequant
out = torch.empty((C, K), dtype=torch.float32, device=dev)
def run(i):
    N.check(L.awqk_dequant(q.data_ptr(), scales.data_ptr(), zp.data_ptr(), C, K, g, out.data_ptr(), st))
t = timeit(run)
res.append(dict(kernel="K4 dequant", us=t * 1e6, hbm_gbs=8 * n / t / 1e9))
print(res[-1], flush=True)
del q, out
# K3 convert
h = torch.empty((C, K), dtype=torch.float16, device=dev)
def run(i):
    N.check(L.awqk_bf16_to_fp16(bufs[i & 1].data_ptr(), h.data_ptr(), n, st))
t = timeit(run)
res.append(dict(kernel="K3 bf16->fp16", us=t * 1e6, hbm_gbs=4 * n / t / 1e9))
print(res[-1], flush=True)
# torch copy for reference (same method as MEASURED_PEAKS)
def run(i):
    h.view(torch.bfloat16).copy_(bufs[i & 1])
t = timeit(run)
res.append(dict(kernel="torch copy_ bf16", us=t * 1e6, hbm_gbs=4 * n / t / 1e9))
print(res[-1], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "probe_k1.json"), "w"), indent=1)

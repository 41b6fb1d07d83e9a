import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "awq-converter_b200"))
import torch
from awq_quantizer import _native as N
L = N.lib(); dev = torch.device("cuda:0")
T, C, K, n = 2048, 4096, 1024, 20
x = torch.randn((T, K), device=dev).to(torch.bfloat16)
w = (torch.randn((C, K), device=dev) * 0.02).to(torch.bfloat16)
sg = torch.rand((n, K), device=dev) + 0.5; rws = torch.empty_like(sg)
dw1 = torch.empty((n, C, K), dtype=torch.bfloat16, device=dev); dw2 = torch.empty_like(dw1)
err = torch.zeros(n, dtype=torch.float64, device=dev)
big = torch.empty(256 << 20, dtype=torch.uint8, device=dev); big2 = torch.empty_like(big)
A, B = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
def gemm(s): N.check(L.awqk_sqerr_gemm(x.data_ptr(), dw1.data_ptr(), T, C, K, n, err.data_ptr(), s.cuda_stream))
def delta(s): N.check(L.awqk_fakequant_delta(w.data_ptr(), N.BF16, C, K, 128, 4, 0, sg.data_ptr(), n, dw2.data_ptr(), s.cuda_stream))
def copy(s):
    with torch.cuda.stream(s): big2.copy_(big)
def run(fa, fb, reps=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); A.wait_event(e0); B.wait_event(e0)
    for _ in range(reps):
        if fa: fa(A)
        if fb: fb(B)
    cur = torch.cuda.current_stream(dev); cur.wait_stream(A); cur.wait_stream(B); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for _ in range(2): gemm(A); delta(B); copy(B)
print("gemm alone   ", run(gemm, None))
print("delta alone  ", run(None, delta))
print("copy alone   ", run(None, copy))
print("gemm || delta", run(gemm, delta))
print("gemm || copy ", run(gemm, copy))
print("delta || copy", run(delta, lambda s: copy(s)))

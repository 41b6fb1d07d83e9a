"""A few stand-alone sqerr GEMM launches (delta operand precomputed in HBM) for ncu capture. ROWS / COLS / TOKENS env."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "awq-converter_b200"))
import torch
from awq_quantizer import _native as N
L = N.lib(); dev = torch.device("cuda:0")
C, K, T = int(os.environ.get("ROWS", "4096")), int(os.environ.get("COLS", "4096")), int(os.environ.get("TOKENS", "2048"))
n = 20
gen = torch.Generator(device=dev).manual_seed(1)
x = torch.randn((T, K), generator=gen, device=dev).to(torch.bfloat16)
dw = (torch.randn((n, C, K), generator=gen, device=dev) * 0.01).to(torch.bfloat16)
err = torch.zeros(n, dtype=torch.float64, device=dev)
st = torch.cuda.current_stream(dev).cuda_stream
for _ in range(3):
    N.check(L.awqk_sqerr_gemm(x.data_ptr(), dw.data_ptr(), T, C, K, n, err.data_ptr(), st))
torch.cuda.synchronize(); print("ok")

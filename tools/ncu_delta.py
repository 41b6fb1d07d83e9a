import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "awq-converter_b200"))
import torch
from awq_quantizer import _native as N
L = N.lib(); dev = torch.device("cuda:0")
C, K, n = 4096, 4096, 20
w = (torch.randn((C, K), device=dev) * 0.02).to(torch.bfloat16)
sg = torch.rand((n, K), device=dev) + 0.5
rws = torch.empty_like(sg)
dw = torch.empty((n, C, K), dtype=torch.bfloat16, device=dev)
st = torch.cuda.current_stream(dev).cuda_stream
for _ in range(3):
    N.check(L.awqk_fakequant_delta(w.data_ptr(), N.BF16, C, K, 128, 4, 0, sg.data_ptr(), n, dw.data_ptr(), st))
torch.cuda.synchronize(); print("ok")

"""One configuration of K1 for ncu capture: python tools/ncu_k1.py [arith] [g] [sym] [mode]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "awq-converter_b200"))
import torch
from awq_quantizer import _native as N
arith = N.ARITH_FP32 if (len(sys.argv) > 1 and sys.argv[1] == "fp32") else N.ARITH_NATIVE
g = int(sys.argv[2]) if len(sys.argv) > 2 else 128
sym = int(sys.argv[3]) if len(sys.argv) > 3 else 0
mode = sys.argv[4] if len(sys.argv) > 4 else "pack"
L = N.lib(); dev = torch.device("cuda:0")
C, K = int(os.environ.get("ROWS", "8192")), 28672
bufs = [(torch.randn((C, K), device=dev, dtype=torch.float32) * 0.02).to(torch.bfloat16) for _ in range(2)]
G = K // g
scales = torch.empty((C, G), dtype=torch.float16, device=dev)
qw = torch.empty((C, K // 8), dtype=torch.int32, device=dev)
qz = torch.empty((C, G // 8), dtype=torch.int32, device=dev)
q = torch.empty((C, K), dtype=torch.int32, device=dev) if mode == "unpacked" else None
zp = torch.empty((C, G), dtype=torch.int32, device=dev) if mode == "unpacked" else None
st = torch.cuda.current_stream(dev).cuda_stream
for i in range(6):
    if mode == "unpacked":
        N.check(L.awqk_group_quant(bufs[i & 1].data_ptr(), N.BF16, C, K, g, 4, sym, arith, q.data_ptr(), None, scales.data_ptr(), zp.data_ptr(), None, None, st))
    else:
        N.check(L.awqk_group_quant(bufs[i & 1].data_ptr(), N.BF16, C, K, g, 4, sym, arith, None, qw.data_ptr(), scales.data_ptr(), None, qz.data_ptr(), None, st))
torch.cuda.synchronize()
print("ok")

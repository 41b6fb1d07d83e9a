"""One column-scaled K1 launch for ncu capture: ROWS / COLS env, argv: [g] [sym] [mode pack|unpacked]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "awq-converter_b200"))
import torch
from awq_quantizer import _native as N
g = int(sys.argv[1]) if len(sys.argv) > 1 else 128
sym = int(sys.argv[2]) if len(sys.argv) > 2 else 0
mode = sys.argv[3] if len(sys.argv) > 3 else "pack"
L = N.lib(); dev = torch.device("cuda:0")
C, K = int(os.environ.get("ROWS", "28672")), int(os.environ.get("COLS", "8192"))
bufs = [(torch.randn((C, K), device=dev, dtype=torch.float32) * 0.02).to(torch.bfloat16) for _ in range(2)]
s = torch.exp(0.5 * torch.randn(K, device=dev)).float()
G = K // g
scales = torch.empty((C, G), dtype=torch.float16, device=dev)
qw = torch.empty((C, K // 8), dtype=torch.int32, device=dev)
qz = torch.empty((C, G // 8), dtype=torch.int32, device=dev)
zp = torch.empty((C, G), dtype=torch.int32, device=dev)
q = torch.empty((C, K), dtype=torch.int32, device=dev) if mode == "unpacked" else None
st = torch.cuda.current_stream(dev).cuda_stream
for i in range(6):
    N.check(L.awqk_group_quant(bufs[i & 1].data_ptr(), N.BF16, C, K, g, 4, sym, N.ARITH_FP32, N.ptr(q), qw.data_ptr(),
                               scales.data_ptr(), zp.data_ptr(), qz.data_ptr(), s.data_ptr(), st))
torch.cuda.synchronize()
print("ok")

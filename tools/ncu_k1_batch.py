"""One column-slab K1 launch over LAYERS Llama-3-8B layers (7 linears each) for ncu capture -- the launch bench.py's
conversion step issues per chunk of 28 searched linears (awqk_group_quant_batch).  env LAYERS (default 4)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "awq-converter_b200"))
import torch
from awq_quantizer import _native as N
dev = torch.device("cuda:0")
st = torch.cuda.current_stream(dev).cuda_stream
layer = [(4096, 4096), (1024, 4096), (1024, 4096), (4096, 4096), (14336, 4096), (14336, 4096), (4096, 14336)]
g = 128
items = []
for C, K in layer * int(os.environ.get("LAYERS", "4")):
    w = (torch.randn((C, K), device=dev, dtype=torch.float32) * 0.02).to(torch.bfloat16)
    s = torch.exp(0.5 * torch.randn(K, device=dev)).float()
    items.append((w, C, K, s, None, torch.empty((C, K // 8), dtype=torch.int32, device=dev),
                  torch.empty((C, K // g), dtype=torch.float16, device=dev), None,
                  torch.empty((C, K // g // 8), dtype=torch.int32, device=dev)))
for _ in range(4):
    N.group_quant_batch(items, N.BF16, g, 4, False, N.ARITH_FP32, st)
torch.cuda.synchronize()
print("ok", sum(C * K for _, C, K, *_ in items), "elements")

"""A few awqk_scale_search calls (fused score kernel + select + column-scaled K1) for ncu capture.
ROWS / COLS / TOKENS env (default 4096 x 4096, T = 2048); argv[1] = number of calls (default 3)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "awq-converter_b200"))
import torch
from awq_quantizer import _native as N
from awq_quantizer.quantization import search as S
dev = torch.device("cuda:0")
C, K, T = int(os.environ.get("ROWS", "4096")), int(os.environ.get("COLS", "4096")), int(os.environ.get("TOKENS", "2048"))
n, calls = 20, int(sys.argv[1]) if len(sys.argv) > 1 else 3
gen = torch.Generator(device=dev).manual_seed(1)
w = (torch.randn((C, K), generator=gen, device=dev) * 0.02).to(torch.bfloat16)
x = (torch.randn((T, K), generator=gen, device=dev) * torch.exp(torch.randn(K, generator=gen, device=dev))).to(torch.bfloat16)
_, grid, xb = S.activation_grid(x, n, N.stream_ptr(dev))
ws = torch.empty(S.workspace_bytes(C, K, T, n)[0], dtype=torch.uint8, device=dev)


class Q:
    group_size, bits = 128, 4


outs = S.alloc_outputs(Q, C, K, dev, pack=True, unpacked=False)
for _ in range(calls):
    r = S.scale_search(w, xb, grid, bits=4, group_size=128, symmetric=False, workspace=ws, outputs=outs)
torch.cuda.synchronize()
print("ok best", int(r["best_idx"]))

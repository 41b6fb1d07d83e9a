"""Timing breakdown of the search leg on the OPT-350m linears: deltas only, GEMMs only, pipeline."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "awq-converter_b200"))
import torch
from awq_quantizer import _native as N, model_shapes as M
from awq_quantizer.quantization.search import SearchPipeline
L = N.lib(); dev = torch.device("cuda:0")
wl = sys.argv[1] if len(sys.argv) > 1 else "opt-350m"
T, n_grid, g = 2048, 20, 128
specs = [(n, s) for n, s, ck in M.workload(wl) if ck is not None]
if len(sys.argv) > 2:
    specs = specs[: int(sys.argv[2])]
shapes = sorted({tuple(s) for _, s in specs})
gen = torch.Generator(device=dev); gen.manual_seed(1)
xs = {K: (torch.randn((T, K), generator=gen, device=dev) * torch.exp(torch.randn(K, generator=gen, device=dev))).to(torch.bfloat16) for K in {s[1] for s in shapes}}
ws = {s: (torch.randn(s, generator=gen, device=dev) * 0.02).to(torch.bfloat16) for s in shapes}
counts = {s: sum(1 for _, t in specs if tuple(t) == s) for s in shapes}
st = torch.cuda.current_stream(dev).cuda_stream
def ev(): return torch.cuda.Event(enable_timing=True)
def timeit(fn):
    fn(); torch.cuda.synchronize()
    a, b = ev(), ev(); t0 = time.perf_counter(); a.record(); fn(); b.record(); host = time.perf_counter() - t0
    torch.cuda.synchronize(); return a.elapsed_time(b), host * 1e3
res = {}
for s in shapes:
    C, K = s
    sg = torch.rand((n_grid, K), device=dev) + 0.5
    rws = torch.empty_like(sg)
    dw = torch.empty((n_grid, C, K), dtype=torch.bfloat16, device=dev)
    err = torch.zeros(n_grid, dtype=torch.float64, device=dev)
    d_ms, _ = timeit(lambda: [N.check(L.awqk_fakequant_delta(ws[s].data_ptr(), N.BF16, C, K, g, 4, 0, sg.data_ptr(), n_grid, dw.data_ptr(), st)) for _ in range(5)])
    g_ms, _ = timeit(lambda: [N.check(L.awqk_sqerr_gemm(xs[K].data_ptr(), dw.data_ptr(), T, C, K, n_grid, err.data_ptr(), st)) for _ in range(5)])
    fl = 2.0 * T * C * K * n_grid
    res[s] = (d_ms / 5, g_ms / 5)
    print(f"{s}: x{counts[s]} delta {d_ms/5*1e3:.1f} us ({n_grid*C*K*2/(d_ms/5*1e-3)/1e9:.0f} GB/s written)  gemm {g_ms/5*1e3:.1f} us ({fl/(g_ms/5*1e-3)/1e12:.0f} TF/s)", flush=True)
    del dw
print("sum delta ms", sum(res[s][0] * counts[s] for s in shapes), "sum gemm ms", sum(res[s][1] * counts[s] for s in shapes))
pipe = SearchPipeline(dev, bits=4, group_size=g, symmetric=False, n_grid=n_grid)
def model():
    for s in shapes:
        for j in range(counts[s]):
            pipe.submit("x", ws[s], xs[s[1]])
    pipe.finish()
ms, host = timeit(model)
print("pipeline ms", ms, "host submit ms", host)

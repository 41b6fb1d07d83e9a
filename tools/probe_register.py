"""Pinned staging memory three ways, 256 MB each: torch.empty(pin_memory=True) (cudaHostAlloc), and a pageable buffer that is
touched by awqk_host_prefault (huge pages, several threads) and then page-locked in place with cudaHostRegister; H2D rate
from each."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "awq-converter_b200"))
import torch
from awq_quantizer import _native as N
L = N.lib()
rt = torch.cuda.cudart()
torch.cuda.init(); torch.zeros(1, device="cuda")
nb = 256 << 20
d = torch.empty(nb, dtype=torch.uint8, device="cuda")
out = []
for rep in range(3):
    t0 = time.perf_counter(); a = torch.empty(nb, dtype=torch.uint8, pin_memory=True); t_alloc = time.perf_counter() - t0
    t0 = time.perf_counter(); b = torch.empty(nb, dtype=torch.uint8); t_empty = time.perf_counter() - t0
    t0 = time.perf_counter(); L.awqk_host_prefault(b.data_ptr(), nb, 0); t_fault = time.perf_counter() - t0
    t0 = time.perf_counter(); rc = rt.cudaHostRegister(b.data_ptr(), nb, 0); t_reg = time.perf_counter() - t0
    pinned = b.is_pinned()
    def h2d(src):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(4): d.copy_(src, non_blocking=True)
        torch.cuda.synchronize(); return 4 * nb / (time.perf_counter() - t0) / 1e9
    r = {"cudaHostAlloc_s": round(t_alloc, 4), "empty_s": round(t_empty, 4), "prefault_s": round(t_fault, 4), "register_s": round(t_reg, 4),
         "register_rc": int(rc), "is_pinned": bool(pinned), "h2d_gbs_alloc": round(h2d(a), 1), "h2d_gbs_registered": round(h2d(b), 1)}
    t0 = time.perf_counter(); rt.cudaHostUnregister(b.data_ptr()); r["unregister_s"] = round(time.perf_counter() - t0, 4)
    out.append(r); print(json.dumps(r), flush=True)
    del a, b
    torch._C._host_emptyCache() if hasattr(torch._C, "_host_emptyCache") else None

"""AWQQuantizer.quantize(tensor) on a large pageable host tensor: routed through the gather pipeline vs the plain
upload -> kernel -> download sequence (the reference's structure, awq.py:402-412)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "awq-converter_b200"))
import torch
from awq_quantizer.quantization import AWQQuantizer
w = (torch.randn(14336, 4096) * 0.02).to(torch.bfloat16)
for pack in (False, True):
    for routed in (True, False):
        qz = AWQQuantizer(bits=4, group_size=128, symmetric=False, device="cuda:0", logger_level="ERROR")
        if not routed:
            qz._PIPELINE_MIN_ELEMS = 1 << 62
        kw = dict(pack=True, keep_unpacked=False) if pack else {}
        qz.quantize(w, **kw); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            r = qz.quantize(w, **kw)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        print(f"quantize(14336x4096 bf16, {'packed' if pack else 'reference layout'}) {'pipelined' if routed else 'plain    '}: "
              f"{dt*1e3:.1f} ms = {w.numel()*2/dt/1e9:.1f} GB/s of BF16 weights", flush=True)

#!/usr/bin/env python
"""awqk_host_copy bandwidth, pageable -> pinned (the staging copy of the upload path), by thread count."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "awq-converter_b200"))
import torch
from awq_quantizer import _native as N
L = N.lib()
nbytes = 1 << 30
src = torch.empty(nbytes, dtype=torch.uint8); src.fill_(1)
dst = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True); dst.fill_(0)
dst2 = torch.empty(nbytes, dtype=torch.uint8); dst2.fill_(0)
out = {"cores": os.cpu_count()}
for name, d in (("pageable_to_pinned", dst), ("pageable_to_pageable", dst2)):
    for th in (1, 2, 4, 8, 12, 16):
        L.awqk_host_copy(d.data_ptr(), src.data_ptr(), nbytes, th)
        t0 = time.perf_counter()
        for _ in range(3):
            L.awqk_host_copy(d.data_ptr(), src.data_ptr(), nbytes, th)
        out[f"{name}_GBps_{th}t"] = round(3 * nbytes / (time.perf_counter() - t0) / 1e9, 1)
t0 = time.perf_counter(); dst2.copy_(src); out["torch_copy_GBps"] = round(nbytes / (time.perf_counter() - t0) / 1e9, 1)
print(json.dumps(out))

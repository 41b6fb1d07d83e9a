#!/usr/bin/env python
"""Cold-start budget of the model-level call in a FRESH process: where the first call's extra time goes.
    python tools/probe_cold.py [--workload llama3-8b] [--layers 4]"""
import argparse, json, os, sys, time
t_proc = time.perf_counter()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "awq-converter_b200"))
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="llama3-8b")
ap.add_argument("--layers", type=int, default=4, help="leading decoder layers (+ embed / lm_head) to convert")
ap.add_argument("--tokens", type=int, default=2048)
ap.add_argument("--no-search", action="store_true")
args = ap.parse_args()
rec = {}
t0 = time.perf_counter(); import torch; rec["import_torch_s"] = time.perf_counter() - t0
t0 = time.perf_counter(); torch.cuda.init(); torch.zeros(1, device="cuda:0"); torch.cuda.synchronize(); rec["cuda_context_s"] = time.perf_counter() - t0
t0 = time.perf_counter()
from awq_quantizer import _native as N, model_shapes as M
from awq_quantizer.quantization import AWQQuantizer
L = N.lib(); rec["import_pkg_dlopen_s"] = time.perf_counter() - t0
dev = torch.device("cuda:0")
specs = [(n, s, ck) for n, s, ck in M.workload(args.workload)
         if M.numel(s) >= 128 and (".layers." not in n or int(n.split(".layers.")[1].split(".")[0]) < args.layers)]
gen = torch.Generator(device=dev)
host_w, acts, xs = {}, {}, {}
import zlib
for n, s, ck in specs:
    gen.manual_seed(zlib.crc32(n.encode()))
    host_w[n] = (torch.randn(s, generator=gen, device=dev) * 0.02).to(torch.bfloat16).cpu()
    if ck is not None and len(s) == 2 and not args.no_search:
        if (ck, s[1]) not in xs:
            xs[(ck, s[1])] = (torch.randn((args.tokens, s[1]), generator=gen, device=dev)).to(torch.bfloat16).cpu().pin_memory()
        acts[n] = xs[(ck, s[1])]
torch.cuda.synchronize()
nbytes = sum(t.numel() * 2 for t in host_w.values())
rec["bf16_GB"] = nbytes / 1e9
# isolated first-use costs
t0 = time.perf_counter(); p = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True); rec["pin_256MB_s"] = time.perf_counter() - t0; del p
t0 = time.perf_counter()
w = torch.zeros((256, 1024), dtype=torch.bfloat16, device=dev); x = torch.zeros((256, 1024), dtype=torch.bfloat16, device=dev)
qz0 = AWQQuantizer(bits=4, group_size=128, symmetric=False, device="cuda:0", logger_level="ERROR")
qz0.quantize(w.cpu(), activations=x.cpu(), pack=True); torch.cuda.synchronize()
rec["first_tiny_search_call_s"] = time.perf_counter() - t0       # kernel module load, func attributes, occupancy query
qz = AWQQuantizer(bits=4, group_size=128, symmetric=False, device="cuda:0", logger_level="ERROR")
times = []
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = qz.quantize_model(host_w, activations=acts or None, pack=True)
    torch.cuda.synchronize(); times.append(time.perf_counter() - t0)
    rec.setdefault("stream_stats", []).append({k: (round(v, 4) if isinstance(v, float) else v)
                                               for k, v in getattr(qz, "last_stream_stats", {}).items()})
    del r
rec["model_call_s"] = [round(t, 4) for t in times]
rec["since_process_start_s"] = time.perf_counter() - t_proc
print(json.dumps(rec))

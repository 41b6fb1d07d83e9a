#!/usr/bin/env python
"""Full-size parity: every tensor of a BASELINE.json model shape quantized on the GPU (K1, native
arithmetic = the reference's own) and compared bit-for-bit with the C restatement of the reference
(oracle/awq_oracle.c, itself pinned to the reference by tests/golden) on the host.

    python tools/verify_model_parity.py --workload llama3-8b [--symmetric] [--group-size 128]
"""
import argparse, json, os, sys, time, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "awq-converter_b200"))
import torch
from awq_quantizer import _native as N, model_shapes as M
from awq_quantizer.quantization import AWQQuantizer
from oracle import awq_oracle as O, c_oracle as CO

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="opt-350m")
ap.add_argument("--group-size", type=int, default=128)
ap.add_argument("--symmetric", action="store_true")
ap.add_argument("--max-tensors", type=int, default=0)
args = ap.parse_args()
dev = torch.device("cuda:0")
g, sym = args.group_size, args.symmetric
qz = AWQQuantizer(bits=4, group_size=g, symmetric=sym, device="cuda:0", logger_level="ERROR")
qmin = -8 if sym else 0
gen = torch.Generator(device=dev)
specs = [(n, s) for n, s, _ in M.workload(args.workload) if M.numel(s) >= 128]
if args.max_tensors:
    specs = specs[: args.max_tensors]
stats = dict(tensors=0, elements=0, groups=0, mismatched_tensors=0, t_gpu=0.0, t_cpu=0.0)
t_all = time.time()
for name, shape in specs:
    gen.manual_seed(zlib.crc32(name.encode()) ^ 0xA11CE)
    w = (torch.randn(shape, generator=gen, device=dev, dtype=torch.float32) * 0.02).to(torch.bfloat16)
    torch.cuda.synchronize(); t0 = time.time()
    r = qz._quantize_device(w, pack=True, unpacked=True)
    torch.cuda.synchronize(); stats["t_gpu"] += time.time() - t0
    host = {k: v.cpu() for k, v in r.items()}
    wc = w.cpu()
    t0 = time.time()
    want = CO.group_quant(wc, 4, g, sym)
    C = shape[0] if len(shape) > 1 else 1
    want_qw = CO.pack_rows(want["tensor_q"].reshape(C, -1), qmin, 4)
    want_qz = CO.pack_rows(want["zero_points"].reshape(C, -1), qmin, 4)
    stats["t_cpu"] += time.time() - t0
    ok = (torch.equal(host["tensor_q"], want["tensor_q"]) and torch.equal(host["zero_points"], want["zero_points"].reshape(host["zero_points"].shape))
          and torch.equal(host["scales"].view(torch.int16), want["scales"].view(torch.int16).reshape(host["scales"].shape))
          and torch.equal(host["qweight"], want_qw) and torch.equal(host["qzeros"], want_qz))
    stats["tensors"] += 1
    stats["elements"] += wc.numel()
    stats["groups"] += want["scales"].numel()
    if not ok:
        stats["mismatched_tensors"] += 1
        print("MISMATCH", name, shape, flush=True)
    del w, r, host, want
stats["wall_s"] = time.time() - t_all
stats.update(workload=args.workload, group_size=g, symmetric=sym, arith="native (bf16)", checker="oracle/awq_oracle.c (OpenMP)")
print(json.dumps(stats))
sys.exit(1 if stats["mismatched_tensors"] else 0)

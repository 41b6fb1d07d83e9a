#!/usr/bin/env python
"""Whole-model AWQ conversion (activation-aware search + final quantize + pack) of a BASELINE.json
model shape with synthetic weights/activations, sharded by tensor over the ranks of one node.

    python tools/run_model.py --workload llama3-8b                      # 1 GPU
    torchrun --nproc-per-node 8 tools/run_model.py --workload llama3-70b

Each rank owns its LPT shard (cost C*K*T for searched linears, bytes otherwise; no data-path
collective), generates its weights on the device (untimed), then times: alpha search for every
linear (SearchPipeline), final group quantization of W * s_best with int4 packing, plain
quantize+pack for the non-linear tensors.  Time = max over ranks (NCCL all_reduce), CUDA events."""
import argparse, json, os, sys, time, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "awq-converter_b200"))
import torch
import torch.distributed as dist
from awq_quantizer import _native as N, model_shapes as M, parallel
from awq_quantizer.quantization.search import SearchPipeline

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="llama3-8b")
ap.add_argument("--tokens", type=int, default=2048)
ap.add_argument("--group-size", type=int, default=128)
ap.add_argument("--n-grid", type=int, default=20)
ap.add_argument("--no-search", action="store_true")
ap.add_argument("--from-host", choices=["pageable", "pinned"], default=None,
                help="end to end through the public API: weights start in host memory, packed results end there")
ap.add_argument("--pin-results", choices=["auto", "0", "1"], default="auto")
args = ap.parse_args()

rank, world = parallel.init_distributed()
local = parallel.rank_info()[2]
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
L = N.lib()
g, T = args.group_size, args.tokens
specs = M.workload(args.workload)
cost = [(n, M.numel(s) * (T if (ck is not None and not args.no_search) else 2)) for n, s, ck in specs]
mine = set(M.partition_lpt(cost, world)[rank])
shard = [(n, s, ck) for n, s, ck in specs if n in mine and M.numel(s) >= 128]

gen = torch.Generator(device=dev)
def synth(name, shape):
    gen.manual_seed(zlib.crc32(name.encode()) ^ 0xA11CE)
    return (torch.randn(shape, generator=gen, device=dev, dtype=torch.float32) * 0.02).to(torch.bfloat16)
weights = {n: synth(n, s) for n, s, _ in shard}
xs = {}
for n, s, ck in shard:
    if ck is not None and s[1] not in xs:
        gen.manual_seed(s[1])
        gain = torch.exp(torch.randn(s[1], generator=gen, device=dev))
        xs[s[1]] = (torch.randn((T, s[1]), generator=gen, device=dev) * gain).to(torch.bfloat16)
if args.from_host:
    # ---- end to end: AWQQuantizer.quantize_model(host tensors, activations=..., pack=True) -----------------
    from awq_quantizer.quantization import AWQQuantizer
    host_w = {}
    for n in list(weights):
        h = weights.pop(n).cpu()
        host_w[n] = h.pin_memory() if args.from_host == "pinned" else h
    host_x = {k: v.cpu().pin_memory() for k, v in xs.items()}
    acts = {} if args.no_search else {n: host_x[s[1]] for n, s, ck in shard if ck is not None}
    del xs
    torch.cuda.empty_cache()
    qz = AWQQuantizer(bits=4, group_size=g, symmetric=False, device=f"cuda:{local}", logger_level="ERROR", n_grid=args.n_grid,
                      pin_results=None if args.pin_results == "auto" else args.pin_results == "1")
    times = []
    for it in range(4):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        res = qz.quantize_model(host_w, activations=acts or None, pack=True)
        torch.cuda.synchronize(dev)
        times.append(time.perf_counter() - t0)
        assert len(res) == len(host_w), (len(res), len(host_w))
        del res
    stat = torch.tensor([min(times[1:]), times[0], sum(M.numel(s) * 2 for _, s, _ in shard), len(acts), len(shard)],
                        device=dev, dtype=torch.float64)
    if world > 1:
        mx = stat[:2].clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stat[2:].clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        stat = torch.cat([mx, sm])
        dist.destroy_process_group()
    if rank == 0:
        best_s, first_s, nbytes, n_lin, n_t = [float(v) for v in stat]
        print(json.dumps({"workload": args.workload, "mode": "end to end from " + args.from_host + " host tensors (public API, packed results on the host)",
                          "n_gpus": world, "tokens": T, "n_grid": args.n_grid, "group_size": g, "s_per_model": best_s,
                          "s_first_call": first_s, "s_calls_rank0": [round(t, 4) for t in times], "searched_linears": int(n_lin), "tensors": int(n_t), "bf16_GB": nbytes / 1e9,
                          "GBps_of_bf16_weights_incl_search": nbytes / best_s / 1e9, "scaling": "strong", "data": "synthetic"}), flush=True)
    sys.exit(0)

# packed outputs (device resident)
outs = {}
for n, s, _ in shard:
    C = s[0] if len(s) > 1 else 1
    K = M.numel(s) // C
    G = -(-K // g)
    outs[n] = (C, K, torch.empty((C, -(-K // 8)), dtype=torch.int32, device=dev), torch.empty((C, G), dtype=torch.float16, device=dev),
               torch.empty((C, G), dtype=torch.int32, device=dev), torch.empty((C, -(-G // 8)), dtype=torch.int32, device=dev))
torch.cuda.synchronize(dev)
if world > 1:
    dist.barrier()

def convert():
    st = torch.cuda.current_stream(dev).cuda_stream
    pipe = SearchPipeline(dev, bits=4, group_size=g, symmetric=False, n_grid=args.n_grid)
    searched = []
    if not args.no_search:
        for n, s, ck in shard:
            if ck is not None:
                pipe.submit(n, weights[n], xs[s[1]])
                searched.append(n)
    best = {name: (mean, idx, s_best) for name, mean, idx, s_best in pipe.finish()}
    for n, s, ck in shard:
        C, K, qw, sc, zp, zq = outs[n]
        if n in best:     # final AWQ quantization of W * s_best (fp32 arithmetic), packed
            N.check(L.awqk_group_quant(weights[n].data_ptr(), N.BF16, C, K, g, 4, 0, N.ARITH_FP32, None, qw.data_ptr(),
                                       sc.data_ptr(), zp.data_ptr(), zq.data_ptr(), best[n][2].contiguous().data_ptr(), st))
        else:             # not a linear: plain quantize + pack (the reference's arithmetic)
            N.check(L.awqk_group_quant(weights[n].data_ptr(), N.BF16, C, K, g, 4, 0, N.ARITH_NATIVE, None, qw.data_ptr(),
                                       sc.data_ptr(), zp.data_ptr(), zq.data_ptr(), None, st))
    return best

convert()                       # warm-up (allocators, lazy module load)
torch.cuda.synchronize(dev)
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
best = convert()
e1.record()
torch.cuda.synchronize(dev)
wall = time.perf_counter() - t0
ms = e0.elapsed_time(e1)
alphas = {n: int(v[1]) for n, v in list(best.items())[:4]}
n_lin = len(best)
flops = sum(2.0 * T * s[0] * s[1] * args.n_grid for n, s, ck in shard if n in best)
nbytes = sum(M.numel(s) * 2 for _, s, _ in shard)
stat = torch.tensor([ms, wall * 1e3, flops, nbytes, n_lin, len(shard)], device=dev, dtype=torch.float64)
if world > 1:
    mx = stat[:2].clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = stat[2:].clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    stat = torch.cat([mx, sm])
    dist.destroy_process_group()
if rank == 0:
    ms, wall_ms, flops, nbytes, n_lin, n_t = [float(v) for v in stat]
    print(json.dumps({"workload": args.workload, "n_gpus": world, "tokens": T, "n_grid": args.n_grid, "group_size": g,
                      "s_per_model": ms * 1e-3, "wall_s_per_model": wall_ms * 1e-3, "searched_linears": int(n_lin),
                      "tensors": int(n_t), "bf16_GB": nbytes / 1e9, "search_tflops_executed": flops / (ms * 1e-3) / 1e12,
                      "GBps_of_bf16_weights_incl_search": nbytes / (ms * 1e-3) / 1e9, "sample_best_idx": alphas,
                      "scaling": "strong", "data": "synthetic"}), flush=True)

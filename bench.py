#!/usr/bin/env python
"""bench.py -- the headline metric of BASELINE.json on B200:

    "quantize+pack GB/s of BF16 weights; end-to-end s/model at 1/2/4/8 B200"

One *step* = one pass of the hot path (int4, g=128 group quantization + nibble pack + packed zero
points + fp16 scales) over every tensor of the workload.  N=1 workload = BASELINE.json configs[1]
(OPT-350m-shaped full convert, synthetic random-init bf16 weights).  With N>1 (torchrun) every
rank owns its LPT share of N model replicas' tensors (weak scaling, no data-path collective; NCCL
only for the barrier / max-over-ranks timing and the final metadata gather).

  value      whole-job GB/s of BF16 weights, inputs resident in HBM, CUDA-event timed
  e2e        same metric through the public API (AWQQuantizer.quantize_model(arena, pack=True)):
             pinned host arena -> chunked H2D -> K1 -> D2H of the packed results, inside the timed region
  roofline   the dominant kernel (K1 group_quant_tma) against the measured HBM peak
  cpu_baseline  the reference's algorithm on the host cores (oracle port; bounded sample)
  search     (when built) the activation-aware alpha search leg: s/model and tensor roofline

`--impl reference` times the reference's own CPU implementation of the path (the group-at-a-time
port of awq.py:332-368 in oracle/awq_oracle.py -- the reference is pure Python and cannot travel to
the GPU box) on a bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "awq-converter_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "quantize+pack GB/s of BF16 weights"
UNIT = "GB/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="opt-350m")
    ap.add_argument("--group-size", type=int, default=128)
    ap.add_argument("--arith", default="native", choices=["native", "fp32"])
    ap.add_argument("--symmetric", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-search", action="store_true")
    ap.add_argument("--search-tokens", type=int, default=2048)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    return ap.parse_args()


# ----------------------------------------------------------------------------- helpers
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i] == "Active"})
        busy = [v for v in sm if v > 0.6 * (mx[0] if mx else 1)] or sm
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": reasons, "samples": len(sm)}


def synth_weight_device(name, shape, dev, torch):
    import zlib
    g = torch.Generator(device=dev)
    g.manual_seed(zlib.crc32(name.encode()) ^ 0xA11CE)
    return (torch.randn(shape, generator=g, device=dev, dtype=torch.float32) * 0.02).to(torch.bfloat16)


def bytes_per_elem(g):
    return 2.0 + 0.5 + 2.0 / g + 0.5 / g      # bf16 in + nibble + fp16 scale/g + 4-bit zero/g


# ----------------------------------------------------------------------------- reference arm
def loop_port_rate(torch, O, w, sym, g, seconds):
    """runs the group-at-a-time port on as many leading rows of `w` as fit in `seconds`"""
    rows = w.shape[0]
    t0 = time.perf_counter()
    O.group_quant_loop(w[:2].contiguous(), 4, g, sym, True)            # calibration (also warms torch)
    per_row = (time.perf_counter() - t0) / 2
    take = max(2, min(rows, int(seconds / max(per_row, 1e-6))))
    sample = w[:take].contiguous()
    t0 = time.perf_counter()
    O.group_quant_loop(sample, 4, g, sym, True)
    dt = time.perf_counter() - t0
    return sample.numel() * 2 / dt / 1e9, take, dt


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle loop port), all host
    threads torch will use; each step = a bounded sample of the workload."""
    import torch
    from awq_quantizer import model_shapes as M
    from oracle import awq_oracle as O
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    specs = [s for s in M.workload(args.workload) if M.numel(s[1]) >= args.group_size]
    name, shape, _ = max(specs, key=lambda s: M.numel(s[1]) if len(s[1]) == 2 else 0)
    torch.manual_seed(0)
    K = shape[1]
    budget = 150.0 / max(1, args.steps + args.warmup)                   # whole run within a few minutes
    w = (torch.randn((4096, K)) * 0.02).to(torch.bfloat16)
    t0 = time.perf_counter()
    O.group_quant_loop(w[:2].contiguous(), 4, args.group_size, args.symmetric, True)
    per_row = (time.perf_counter() - t0) / 2
    rows = max(2, min(4096, int(budget / max(per_row, 1e-6))))
    sample = w[:rows].contiguous()
    for _ in range(args.warmup):
        O.group_quant_loop(sample, 4, args.group_size, args.symmetric, True)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.group_quant_loop(sample, 4, args.group_size, args.symmetric, True)
    dt = (time.perf_counter() - t0) / args.steps
    val = sample.numel() * 2 / dt / 1e9
    sample_desc = (f"{rows} rows x {K} of a {args.workload}-shaped bf16 linear per step "
                   f"({sample.numel() // args.group_size} groups), group-at-a-time port of awq.py:332-368")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16 (reference arithmetic dtype)", "data": "synthetic",
        "config": {"workload": f"{args.workload}-shaped, int4 g{args.group_size} "
                               f"{'symmetric' if args.symmetric else 'asymmetric'}", "sample": sample_desc},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample_desc},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------- native arm
def run_native(args):
    import torch
    import torch.distributed as dist
    from awq_quantizer import _native as N
    from awq_quantizer import model_shapes as M
    from awq_quantizer.quantization import AWQQuantizer
    from awq_quantizer.quantization.arena import HostArena, arena_eligible

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    from awq_quantizer import parallel
    numa_node = parallel.bind_to_gpu_numa(local) if world > 1 else None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = N.lib()
    g, bits, sym = args.group_size, 4, args.symmetric
    per = 32 // bits
    hbm_peak, tf_peak, peak_kind = measured_peaks()

    # ---- this rank's share: LPT partition of `world` replicas of the workload (weak scaling) ----
    specs = M.workload(args.workload)
    pool = [(f"r{r}/{name}", shape) for r in range(world) for name, shape, _ in specs]
    bins = M.partition_lpt([(n, M.numel(s) * 2) for n, s in pool], world)
    mine = set(bins[rank])
    shapes = {n: s for n, s in pool if n in mine}
    flat = {n: s for n, s in shapes.items() if arena_eligible(s, torch.bfloat16, g, bits)}
    single = {n: s for n, s in shapes.items() if n not in flat and M.numel(s) >= 128}   # CLI drops numel < 128 (main.py:250)
    payload_elems = sum(M.numel(s) for s in flat.values()) + sum(M.numel(s) for s in single.values())
    payload_bytes = 2 * payload_elems

    # ---- synthetic inputs: generated on the device, one D2H into the pinned host arena (untimed) ----
    arena = HostArena({n: (tuple(s), torch.bfloat16) for n, s in flat.items()})
    d_arena = torch.zeros_like(arena.buffers[torch.bfloat16], device=dev)
    for name, off, n in arena.layout[torch.bfloat16]:
        d_arena[off:off + n] = synth_weight_device(name, flat[name], dev, torch).reshape(-1)
    arena.buffers[torch.bfloat16].copy_(d_arena)
    d_single = {n: synth_weight_device(n, s, dev, torch) for n, s in single.items()}
    h_single = {n: t.cpu().pin_memory() for n, t in d_single.items()}
    n_arena = d_arena.numel()

    # ---- device-resident leg ----------------------------------------------------------------------
    d_q = torch.empty(n_arena // per, dtype=torch.int32, device=dev)
    d_s = torch.empty(n_arena // g, dtype=torch.float16, device=dev)
    d_zq = torch.empty(n_arena // g // per, dtype=torch.int32, device=dev)
    single_out = {}
    for n, t in d_single.items():
        C = t.shape[0] if t.dim() > 1 else 1
        K = t.numel() // C
        G = -(-K // g)
        single_out[n] = (C, K, torch.empty((C, -(-K // per)), dtype=torch.int32, device=dev),
                         torch.empty((C, G), dtype=torch.float16, device=dev),
                         torch.empty((C, G), dtype=torch.int32, device=dev),
                         torch.empty((C, -(-G // per)), dtype=torch.int32, device=dev))
    arith = N.ARITH_FP32 if args.arith == "fp32" else N.ARITH_NATIVE
    # a row-mode tensor is one K1 launch when its rows are 1 / 2 / 4 groups (zero words packed in the kernel),
    # else K1 + pack_zeros_rows
    launches_per_step = 1 + sum(1 if (bits == 4 and v[1] // g in (1, 2, 4)) else 2 for v in single_out.values())

    def k1_arena():
        st = torch.cuda.current_stream(dev).cuda_stream
        N.check(L.awqk_group_quant(d_arena.data_ptr(), N.BF16, 1, n_arena, g, bits, int(sym), arith, None,
                                   d_q.data_ptr(), d_s.data_ptr(), None, d_zq.data_ptr(), None, st))

    def step_launches():
        # the (small) row-mode tensors first, then the arena: one stream, no fork / join.  (A forked stream
        # buys nothing: the persistent arena kernel fills every SM, so the small kernels would only run at its tail.)
        st = torch.cuda.current_stream(dev).cuda_stream
        for n, t in d_single.items():
            C, K, qw, sc, zp, zq = single_out[n]
            N.check(L.awqk_group_quant(t.data_ptr(), N.BF16, C, K, g, bits, int(sym), arith, None, qw.data_ptr(),
                                       sc.data_ptr(), None if (bits == 4 and K // g in (1, 2, 4)) else zp.data_ptr(), zq.data_ptr(), None, st))
        k1_arena()

    # one pass = one CUDA graph launch (the per-tensor launches are captured once, replayed per step)
    step_device, graph_mode = step_launches, "direct launches"
    try:
        step_launches()
        torch.cuda.synchronize(dev)
        side = torch.cuda.Stream(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            step_launches()
        graph.replay()
        torch.cuda.synchronize(dev)
        step_device, graph_mode = graph.replay, "CUDA graph replay"
    except Exception as e:      # capture is an optimisation of the launch path only
        graph_mode = f"direct launches (graph capture failed: {str(e)[:80]})"
        torch.cuda.synchronize(dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    clocks = ClockSampler(local)
    clocks.start()
    ms_step = timed(step_device, args.steps, max(3, args.warmup))
    ms_k1 = timed(k1_arena, args.steps, 3)                       # the dominant kernel alone (roofline)
    # (the clock sampler keeps running through the e2e and search legs: all of them are timed regions)

    total_payload = payload_bytes
    if world > 1:
        t = torch.tensor([payload_bytes], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        total_payload = float(t.item())
    value = total_payload / (ms_step * 1e-3) / 1e9

    arena_payload_elems = sum(M.numel(s) for s in flat.values())
    achieved = bytes_per_elem(g) * arena_payload_elems / (ms_k1 * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic = tj.get(f"{args.workload}/g{g}/{args.arith}")
    roofline = {"bound": "hbm", "kernel": "group_quant_tma (K1 v2)", "achieved": achieved, "peak": hbm_peak,
                "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_kind,
                "us_per_launch": ms_k1 * 1e3, "algorithmic_bytes_per_launch": bytes_per_elem(g) * arena_payload_elems}

    # ---- e2e leg: public API, pinned host arena in, packed host results out --------------------------
    qz = AWQQuantizer(bits=bits, group_size=g, symmetric=sym, device=f"cuda:{local}", logger_level="ERROR",
                      arith=args.arith)

    def step_e2e():
        r = qz.quantize_model(arena, pack=True)
        r.update(qz.quantize_model(h_single, pack=True))       # rows that are not whole packed-zero words
        return r

    for _ in range(2):
        res = step_e2e()
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = step_e2e()
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e_val = total_payload / (dt / e2e_steps) / 1e9
    # the unchanged reference call -- quantize_model(dict) -> int32 tensor_q / fp16 scales / int32 zero points --
    # for the record (D2H of 4 B per element makes it PCIe-bound at ~1/4 of the packed path)
    host_dict = {n: arena.views[n].clone() for n in flat}      # ordinary pageable tensors, as a loader returns them
    host_dict.update({n: t.clone() for n, t in h_single.items()})
    # ... and the packed form of the same call: first call of a fresh quantizer (pageable results through the
    # native gather pipeline's bounded pinned ring -- what a one-shot conversion sees), then warm calls
    qz2 = AWQQuantizer(bits=bits, group_size=g, symmetric=sym, device=f"cuda:{local}", logger_level="ERROR",
                       arith=args.arith)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    qz2.quantize_model(host_dict, pack=True)
    torch.cuda.synchronize(dev)
    dt_dict_first = time.perf_counter() - t0
    qz2.quantize_model(host_dict, pack=True)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(3):
        qz2.quantize_model(host_dict, pack=True)
    torch.cuda.synchronize(dev)
    dt_dict_pack = (time.perf_counter() - t0) / 3
    qz.quantize_model(host_dict)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(3):
        qz.quantize_model(host_dict)
    torch.cuda.synchronize(dev)
    dt_ref_layout = (time.perf_counter() - t0) / 3
    h2d = arena.nbytes() + sum(t.numel() * 2 for t in h_single.values())
    d2h = (n_arena // per) * 4 + (n_arena // g) * 2 + (n_arena // g // per) * 4
    d2h += sum(sum(v.numel() * v.element_size() for k, v in res[n].items() if v.dim() > 0) for n in h_single)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.arith == "native" else "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}-shaped full convert x{world} (LPT over ranks), int4 g{g} "
                               f"{'symmetric' if sym else 'asymmetric'}, arith={args.arith}",
                   "tensors_per_rank": len(shapes), "params_per_rank": payload_elems,
                   "l2": "inputs larger than L2 (arena %.0f MB per pass)" % (n_arena * 2 / 1e6),
                   "parallelism": f"tensor-sharded x{world}, no data-path collective", "launch": graph_mode,
                   "numa_node_rank0": numa_node},
        "s_per_model": ms_step * 1e-3, "roofline": roofline,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "s_per_model": dt / e2e_steps, "api": "AWQQuantizer.quantize_model(HostArena, pack=True)",
                "from_pageable_dict": {"value": payload_bytes / dt_dict_pack / 1e9, "unit": UNIT, "s_per_model": dt_dict_pack,
                                       "first_call_s": dt_dict_first,
                                       "api": "AWQQuantizer.quantize_model(dict of pageable tensors, pack=True)  (per rank)"},
                "reference_layout": {"value": payload_bytes / dt_ref_layout / 1e9, "unit": UNIT, "s_per_model": dt_ref_layout,
                                     "api": "AWQQuantizer.quantize_model(dict)  (unchanged reference call; per rank)"}},
        "gpu_launches": launches_per_step * args.steps,
    }

    # ---- activation-aware search leg (K2), when built ----------------------------------------------
    if not args.no_search:
        try:
            from awq_quantizer.quantization import search as S
            line["search"] = S.bench_leg(args, dev, world, rank, tf_peak, peak_kind)
        except ImportError:
            line["search"] = None

    line["clocks"] = clocks.stop()

    # ---- CPU baseline: the oracle on this box's host cores (rank 0, N=1 only) -----------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import awq_oracle as O
        name, shape = max(flat.items(), key=lambda kv: M.numel(kv[1]) if len(kv[1]) == 2 and kv[1][1] >= 1024 else 0)
        w = arena.views[name][:4096].clone()
        gbs, rows, secs = loop_port_rate(torch, O, w, sym, g, args.cpu_seconds)
        cb = {"value": gbs, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
              "sample": f"first {rows} rows of {name.split('/', 1)[1]} {tuple(shape)} bf16, {secs:.1f}s of the "
                        f"group-at-a-time port (awq.py:332-368)"}
        try:
            from oracle import c_oracle as CO
            big = arena.views[name]
            t0 = time.perf_counter()
            CO.group_quant(big, 4, g, sym, threads=os.cpu_count())
            cb["c_restatement_all_cores"] = {"value": big.numel() * 2 / (time.perf_counter() - t0) / 1e9, "unit": UNIT,
                                             "cores": os.cpu_count(), "sample": f"{name.split('/', 1)[1]} whole tensor"}
        except Exception as e:   # C oracle is optional test infrastructure
            cb["c_restatement_all_cores"] = {"error": str(e)[:100]}
        line["cpu_baseline"] = cb

    if world > 1:
        meta = [None] * world
        dist.all_gather_object(meta, {"rank": rank, "tensors": len(shapes), "bytes": payload_bytes})
        if rank == 0:
            line["config"]["per_rank"] = meta
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)
    return 0


def main():
    args = parse()
    # stdout carries exactly ONE line (the JSON record): libraries that write to file descriptor 1 behind Python's
    # back (NCCL prints its version banner there when NCCL_DEBUG is set in the environment) are sent to stderr.
    sys.stdout.flush()
    real_out = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_out, "w", buffering=1)
    if args.impl == "reference":
        return run_reference(args)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())

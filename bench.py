#!/usr/bin/env python
"""bench.py -- the headline metric of BASELINE.json on B200:

    "quantize+pack GB/s of BF16 weights; end-to-end s/model at 1/2/4/8 B200"

One *step* = one full AWQ conversion of the workload's tensors that this job owns: for every nn.Linear weight the
20-point activation-aware alpha search (column statistic -> scale grid -> fused fake-quant producer + tcgen05 GEMM
scores -> device argmin) and the final column-scaled int4 g128 group quantization + nibble pack + packed zeros +
fp16 scales; for every other tensor the plain group quantization + pack.  All of it is inside `value`.

Workloads (BASELINE.json configs):  --gpus 1 -> Llama-3-8B shape (configs[2]); --gpus 2/4/8 -> Llama-3-70B shape
(configs[3]) partitioned over the ranks by tensor (LPT, the reference's own never-called partition_tensors,
main.py:395-427) with NO data-path collective -- strong scaling over 2/4/8.  `--workload` overrides.

  value        whole-job GB/s of BF16 weights (model bytes / max-over-ranks step time), inputs resident in HBM,
               CUDA events around >= 1 s of back-to-back steps
  e2e          the same conversion through the call a user makes -- AWQQuantizer.quantize_model(dict of ordinary
               (pageable) host tensors, activations=..., pack=True) -- host->device and device->host inside
  roofline     the dominant kernel (search_fused_kernel: tensor-core bound) against the measured bf16 peak
  pack         the K1 launches of the same model alone (HBM bound): burst and sustained, against the measured HBM peak
  cpu_baseline the reference's own Python loop (oracle/_ref when vendored by build(), else the oracle port) on a
               bounded sample, host cores stated

`--impl reference` times the reference's CPU implementation of the path on a bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import zlib

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "awq-converter_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "quantize+pack GB/s of BF16 weights"
UNIT = "GB/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="auto", help="auto: llama3-8b on 1 GPU, llama3-70b on 2/4/8")
    ap.add_argument("--group-size", type=int, default=128)
    ap.add_argument("--symmetric", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-search", action="store_true", help="plain quantize+pack of every tensor (no activations)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--search-tokens", type=int, default=2048)
    ap.add_argument("--n-grid", type=int, default=20)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--profile", action="store_true",
                    help="for ncu launch lists: one warm-up step, --steps timed steps, no other leg (not a bench number)")
    return ap.parse_args()


def pick_workload(args, world):
    if args.workload != "auto":
        return args.workload
    return "llama3-8b" if world == 1 else "llama3-70b"


# ----------------------------------------------------------------------------- helpers
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        sus = float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0)))
        return {"hbm": float(d["hbm_gbs"]), "tf_sustained": sus,
                "tf_burst": float(d.get("bf16_tflops_burst", d.get("bf16_tflops", sus))),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm": 6650.0, "tf_sustained": 1590.0, "tf_burst": 1590.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed regions (B200_PROFILING.md)."""

    def __init__(self, index: int, period_ms: int = 200):
        self.index, self.period_ms = index, period_ms
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", str(self.period_ms)],
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
        time.sleep(0.25)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        pw = sorted(float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "", 1).isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i] == "Active"})
        busy = [v for v in sm if v > 0.6 * (mx[0] if mx else 1)] or sm
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": reasons, "samples": len(sm), "power_w_median": pw[len(pw) // 2] if pw else None}


def numa_nodes():
    try:
        return len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()])
    except OSError:
        return None


def bytes_per_elem(g):
    return 2.0 + 0.5 + 2.0 / g + 0.5 / g      # bf16 in + nibble + fp16 scale/g + 4-bit zero/g


def seed_of(name: str) -> int:
    return zlib.crc32(name.encode()) ^ 0xA11CE


# ----------------------------------------------------------------------------- the reference on the host cores
def load_reference_quantizer():
    """the UNMODIFIED reference package, vendored by __graft_entry__.build() into the git-ignored oracle/_ref/
    (it travels to the GPU box with the snapshot; /root/reference does not exist there).  None when absent."""
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "awq_quantizer")):
        return None
    import importlib.util
    path = os.path.join(ref_dir, "awq_quantizer", "quantization", "awq.py")
    # loaded under a private name so that it can never shadow the product package `awq_quantizer`
    import types
    pkg_names = ["_awq_ref", "_awq_ref.utils", "_awq_ref.quantization"]
    for nm in pkg_names:
        if nm not in sys.modules:
            m = types.ModuleType(nm)
            m.__path__ = [os.path.join(ref_dir, "awq_quantizer", *nm.split(".")[1:])]
            sys.modules[nm] = m
    try:
        for sub in ("utils.logger", "utils.tensor_utils", "quantization.awq"):
            nm = "_awq_ref." + sub
            if nm in sys.modules:
                continue
            spec = importlib.util.spec_from_file_location(nm, os.path.join(ref_dir, "awq_quantizer", *sub.split(".")) + ".py")
            mod = importlib.util.module_from_spec(spec)
            sys.modules[nm] = mod
            spec.loader.exec_module(mod)
        return sys.modules["_awq_ref.quantization.awq"].AWQQuantizer
    except Exception as e:                       # an unexpected reference layout: report the port instead
        sys.stderr.write(f"[bench] vendored reference not importable ({e}); using the oracle port\n")
        return None


class CpuReference:
    """quantize(w) of the reference on the host cores: the real class when vendored, else the group-at-a-time port"""

    def __init__(self, g, sym):
        import torch
        self.torch, self.g, self.sym = torch, g, sym
        cls = load_reference_quantizer()
        if cls is not None:
            self.kind = "reference"
            self.qz = cls(bits=4, group_size=g, symmetric=sym, device="cpu", logger_level="ERROR")
            self.what = "unmodified AWQQuantizer(device='cpu').quantize (awq.py:376) from oracle/_ref"
        else:
            from oracle import awq_oracle as O
            self.kind, self.O = "port", O
            self.what = "group-at-a-time port of awq.py:332-368 (oracle/awq_oracle.py::group_quant_loop)"

    def run(self, w):
        if self.kind == "reference":
            return self.qz.quantize(w)
        return self.O.group_quant_loop(w, 4, self.g, self.sym, True)

    def rows_for(self, w, seconds):
        self.run(w[:2].contiguous())                                    # warms torch / the logger
        t0 = time.perf_counter()
        self.run(w[:8].contiguous())
        per_row = (time.perf_counter() - t0) / 8
        return max(2, min(w.shape[0], int(seconds / max(per_row, 1e-6))))


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path, all host threads torch will use; each step
    = a bounded sample (leading rows of the workload's most common linear shape).  The reference has no scale
    search: its step is the group quantization alone -- less work than the native arm's."""
    import torch
    from awq_quantizer import model_shapes as M
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return 0
    workload = pick_workload(args, world)
    specs = [s for s in M.workload(workload) if len(s[1]) == 2 and s[2] is not None]
    name, shape, _ = max(specs, key=lambda s: M.numel(s[1]))
    torch.manual_seed(0)
    K = shape[1]
    ref = CpuReference(args.group_size, args.symmetric)
    budget = 150.0 / max(1, args.steps + args.warmup)                   # whole run within a few minutes
    w = (torch.randn((min(shape[0], 16384), K)) * 0.02).to(torch.bfloat16)
    rows = ref.rows_for(w, budget)
    sample = w[:rows].contiguous()
    for _ in range(args.warmup):
        ref.run(sample)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.run(sample)
    dt = (time.perf_counter() - t0) / args.steps
    val = sample.numel() * 2 / dt / 1e9
    sample_desc = (f"{rows} rows x {K} of {name.split('.')[-2]} {tuple(shape)} of the {workload} shape per step "
                   f"({sample.numel() // args.group_size} groups); {ref.what}")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "bf16 (reference arithmetic dtype)",
        "data": "synthetic",
        "config": {"workload": f"{workload}-shaped, int4 g{args.group_size} "
                               f"{'symmetric' if args.symmetric else 'asymmetric'}", "sample": sample_desc,
                   "note": "group quantization only: the reference has no activation-aware search to time"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": ref.kind,
                         "sample": sample_desc},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------- native arm
class DeviceModel:
    """This rank's shard of the workload, resident in HBM, and the launch sequences of one conversion.
    Searched linears: one awqk_scale_search call each (scores, argmin, winning scales) and ONE awqk_group_quant_batch
    call per chunk of linears for the final column-scaled pass, as quantization/search.py::SearchPipeline does per
    wave.  Every other tensor sits in one tile-aligned device arena per shard (the layout of quantization/arena.py
    and of the gather pipeline's chunks) and is quantized by ONE flat K1 launch."""

    FINAL_CHUNK = 28          # linears per awqk_group_quant_batch call (four Llama layers, ~0.9 GB of weights)
    TILE = 8192               # K1's CTA tile: arena slots are aligned to it

    def __init__(self, torch, N, M, shard, dev, *, g, sym, T, n_grid, search):
        self.torch, self.N, self.L, self.dev = torch, N, N.lib(), dev
        self.g, self.sym, self.T, self.n_grid = g, sym, T, n_grid
        gen = torch.Generator(device=dev)
        self.items = []          # (name, C, K, calib key or None)
        self.w, self.out, self.x, self.grid = {}, {}, {}, {}
        ws_bytes = 256
        plain = []
        for name, shape, ck in shard:
            C = shape[0] if len(shape) > 1 else 1
            K = M.numel(shape) // C
            ck = ck if (search and ck is not None and len(shape) == 2) else None
            self.items.append((name, C, K, ck))
            if ck is None:
                plain.append((name, shape, C, K))
                continue
            G = -(-K // g)
            gen.manual_seed(seed_of(name))
            self.w[name] = (torch.randn(shape, generator=gen, device=dev, dtype=torch.float32) * 0.02).to(torch.bfloat16)
            packed_zeros_in_kernel = (G % 8 == 0) or G in (1, 2, 4)
            self.out[name] = {"qweight": torch.empty((C, -(-K // 8)), dtype=torch.int32, device=dev),
                              "scales": torch.empty((C, G), dtype=torch.float16, device=dev),
                              "zero_points": None if packed_zeros_in_kernel else torch.empty((C, G), dtype=torch.int32, device=dev),
                              "qzeros": torch.empty((C, -(-G // 8)), dtype=torch.int32, device=dev)}
            key = (ck, K)
            if key not in self.x:
                gen.manual_seed(seed_of(f"x/{ck}/{K}"))
                gain = torch.exp(torch.randn(K, generator=gen, device=dev))
                self.x[key] = (torch.randn((T, K), generator=gen, device=dev) * gain).to(torch.bfloat16)
                self.grid[key] = (torch.empty(K, dtype=torch.float64, device=dev),
                                  torch.empty((n_grid, K), dtype=torch.float32, device=dev),
                                  torch.empty(2 * n_grid, dtype=torch.float32, device=dev))
            import ctypes
            ws_bytes = max(ws_bytes, int(self.L.awqk_workspace_bytes(C, K, T, n_grid, 1, ctypes.byref(ctypes.c_size_t(0)))))
        # ---- plain tensors: one arena (whole groups only; anything else keeps its own launch) ----
        self.arena = None
        self.loose = []
        slots, off = [], 0
        for name, shape, C, K in plain:
            if K % (8 * g) == 0:                      # rows of whole packed zero-point words: flat layout == row layout
                slots.append((name, shape, C, K, off))
                off += -(-(C * K) // self.TILE) * self.TILE
            else:
                self.loose.append((name, C, K))
        if slots:
            a = torch.zeros(off, dtype=torch.bfloat16, device=dev)
            ao = {"qweight": torch.empty(off // 8, dtype=torch.int32, device=dev),
                  "scales": torch.empty(off // g, dtype=torch.float16, device=dev),
                  "qzeros": torch.empty(off // g // 8, dtype=torch.int32, device=dev)}
            for name, shape, C, K, o in slots:
                gen.manual_seed(seed_of(name))
                a[o:o + C * K].view(shape).copy_((torch.randn(shape, generator=gen, device=dev, dtype=torch.float32) * 0.02).to(torch.bfloat16))
                self.w[name] = a[o:o + C * K].view(shape)
                G = K // g
                self.out[name] = {"qweight": ao["qweight"][o // 8:(o + C * K) // 8].view(C, K // 8),
                                  "scales": ao["scales"][o // g:(o + C * K) // g].view(C, G), "zero_points": None,
                                  "qzeros": ao["qzeros"][o // g // 8:(o + C * K) // g // 8]}
            self.arena = (a, ao, off)
        for name, C, K in self.loose:
            G = -(-K // g)
            gen.manual_seed(seed_of(name))
            self.w[name] = (torch.randn((C, K), generator=gen, device=dev, dtype=torch.float32) * 0.02).to(torch.bfloat16)
            self.out[name] = {"qweight": torch.empty((C, -(-K // 8)), dtype=torch.int32, device=dev),
                              "scales": torch.empty((C, G), dtype=torch.float16, device=dev),
                              "zero_points": torch.empty((C, G), dtype=torch.int32, device=dev),
                              "qzeros": torch.empty((C, -(-G // 8)), dtype=torch.int32, device=dev)}
        self.searched = [it for it in self.items if it[3] is not None]
        self.plain = [it for it in self.items if it[3] is None]
        self.workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        self.sel = {n: (torch.empty(n_grid, dtype=torch.float64, device=dev), torch.empty((), dtype=torch.int32, device=dev),
                        torch.empty(K, dtype=torch.float32, device=dev)) for n, _, K, _ in self.searched}
        self.elems = sum(C * K for _, C, K, _ in self.items)
        self.search_flops = sum(2.0 * T * C * K * n_grid for _, C, K, _ in self.searched)
        self.final_calls = -(-len(self.searched) // self.FINAL_CHUNK)
        # kernels of ours per conversion: grids (colsum + 3 alpha-grid kernels), per linear fused scores + select,
        # one column-slab K1 launch per chunk of linears, one flat K1 launch for the arena (+ the loose tensors)
        self.k1_launches = self.final_calls + (1 if self.arena else 0) + 2 * len(self.loose)
        self.launches = 4 * len(self.x) + 2 * len(self.searched) + self.k1_launches

    def _st(self):
        return self.torch.cuda.current_stream(self.dev).cuda_stream

    def grids(self):
        N, L, st = self.N, self.L, self._st()
        for key, x in self.x.items():
            colsum, s_grid, mnmx = self.grid[key]
            colsum.zero_()
            N.check(L.awqk_abs_colsum(x.data_ptr(), N.BF16, x.shape[0], x.shape[1], colsum.data_ptr(), st), "awqk_abs_colsum")
            N.check(L.awqk_alpha_grid(colsum.data_ptr(), x.shape[0], x.shape[1], self.n_grid, s_grid.data_ptr(),
                                      mnmx.data_ptr(), st), "awqk_alpha_grid")

    def finals(self, items):
        """the final column-scaled pass of a chunk of searched linears: one awqk_group_quant_batch call"""
        N = self.N
        batch = []
        for name, C, K, _ in items:
            o = self.out[name]
            batch.append((self.w[name], C, K, self.sel[name][2], None, o["qweight"], o["scales"], o["zero_points"], o["qzeros"]))
        N.group_quant_batch(batch, N.BF16, self.g, 4, self.sym, N.ARITH_FP32, self._st())

    def search(self, final: bool):
        """awqk_scale_search per linear: scores + argmin + winning scales; the final pass chunk by chunk"""
        N, L, st = self.N, self.L, self._st()
        ws = self.workspace
        for i, (name, C, K, ck) in enumerate(self.searched):
            x = self.x[(ck, K)]
            s_grid = self.grid[(ck, K)][1]
            err, best, s_best = self.sel[name]
            N.check(L.awqk_scale_search(
                self.w[name].data_ptr(), N.BF16, C, K, x.data_ptr(), self.T, s_grid.data_ptr(), self.n_grid, self.g, 4,
                int(self.sym), err.data_ptr(), best.data_ptr(), s_best.data_ptr(), None, None, None, None, None,
                ws.data_ptr(), ws.numel(), st), "awqk_scale_search")
            if final and ((i + 1) % self.FINAL_CHUNK == 0 or i + 1 == len(self.searched)):
                self.finals(self.searched[i - i % self.FINAL_CHUNK:i + 1])

    def k1_plain(self):
        N, L, st = self.N, self.L, self._st()
        if self.arena:
            a, ao, n = self.arena
            N.check(L.awqk_group_quant(a.data_ptr(), N.BF16, 1, n, self.g, 4, int(self.sym), N.ARITH_NATIVE, None,
                                       ao["qweight"].data_ptr(), ao["scales"].data_ptr(), None, ao["qzeros"].data_ptr(),
                                       None, st), "awqk_group_quant (arena)")
        for name, C, K in self.loose:
            o = self.out[name]
            N.check(L.awqk_group_quant(self.w[name].data_ptr(), N.BF16, C, K, self.g, 4, int(self.sym), N.ARITH_NATIVE, None,
                                       o["qweight"].data_ptr(), o["scales"].data_ptr(), N.ptr(o["zero_points"]),
                                       o["qzeros"].data_ptr(), None, st), "awqk_group_quant")

    def convert(self):                       # ONE step
        self.grids()
        self.search(final=True)
        self.k1_plain()

    def scores_only(self):                   # the dominant kernel's launches alone (roofline)
        self.search(final=False)

    def pack_only(self):                     # every K1 launch of the conversion alone (HBM roofline)
        for i in range(0, len(self.searched), self.FINAL_CHUNK):
            self.finals(self.searched[i:i + self.FINAL_CHUNK])
        self.k1_plain()


def run_native(args):
    import torch
    import torch.distributed as dist
    from awq_quantizer import _native as N
    from awq_quantizer import model_shapes as M
    from awq_quantizer import parallel
    from awq_quantizer.quantization import AWQQuantizer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    numa_node = parallel.bind_to_gpu_numa(local) if world > 1 else None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    g, sym, T, n_grid = args.group_size, args.symmetric, args.search_tokens, args.n_grid
    search = not args.no_search
    peaks = measured_peaks()
    workload = pick_workload(args, world)

    # ---- this rank's share: LPT by cost (C*K*T for searched linears, bytes otherwise), no data-path collective ----
    specs = [(n, s, ck) for n, s, ck in M.workload(workload) if M.numel(s) >= 128]      # CLI drops numel < 128 (main.py:250)
    cost = [(n, M.numel(s) * (T if (search and ck is not None and len(s) == 2) else 2)) for n, s, ck in specs]
    mine = set(M.partition_lpt(cost, world)[rank])
    shard = [(n, s, ck) for n, s, ck in specs if n in mine]
    model = DeviceModel(torch, N, M, shard, dev, g=g, sym=sym, T=T, n_grid=n_grid, search=search)
    payload_bytes = 2 * model.elems
    total_bytes = 2 * sum(M.numel(s) for _, s, _ in specs)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def allmax(v):
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps, warmup):
        """(max-over-ranks ms per step, this rank's ms per step)"""
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        mine_ms = e0.elapsed_time(e1) / steps
        return allmax(mine_ms), mine_ms

    # ---- device-resident leg: K whole conversions back to back ------------------------------------------------
    warm = 1 if args.profile else max(3, args.warmup)
    if args.profile:
        torch.cuda.synchronize(dev)
        torch.cuda.cudart().cudaProfilerStart()     # ncu --profile-from-start off: skip the synthetic-data kernels
    for _ in range(warm):
        model.convert()
    clocks = ClockSampler(local)                   # sampled over the timed region of the headline number only
    clocks.start()
    ms_step, ms_step_mine = timed(model.convert, args.steps, 0)
    clocks_step = clocks.stop()
    value = total_bytes / (ms_step * 1e-3) / 1e9
    timed_region_s = ms_step * args.steps * 1e-3

    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_run": True, "ms_per_step": ms_step, "gpu_launches": model.launches * args.steps}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- dominant kernel: the fused score kernel, same launches as in the step, alone, back to back (sustained) ----
    roofline, ms_scores_mine = None, 0.0
    if model.searched:
        model.grids()
        ms_scores, ms_scores_mine = timed(model.scores_only, max(1, min(args.steps, 3)), 1)
        flops = model.search_flops
        achieved = flops / (ms_scores * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "search_fused_kernel (fake-quant producer warps + tcgen05 cta_group::2 GEMM)",
                    "achieved": achieved, "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["tf_sustained"],
                    "frac_of_burst_peak": achieved / peaks["tf_burst"], "traffic": None, "peak_source": peaks["source"],
                    "peak_kind": "sustained (kernel timed inside %.2f s of back-to-back launches)" % (ms_scores * 1e-3),
                    "launches": len(model.searched), "us_per_launch": ms_scores * 1e3 / len(model.searched),
                    "algorithmic_flops_per_launch": flops / len(model.searched),
                    "flops_formula": "2*T*C*K*n_grid (delta form; SURVEY's 2*T*C*K*(n_grid+1) gives %.1f TFLOP/s)"
                                     % (achieved * (n_grid + 1) / n_grid),
                    "share_of_step": ms_scores / ms_step_mine}

    # ---- pack kernels alone (HBM bound): one pass = burst, >= 1 s back to back = sustained ---------------------
    alg_bytes = bytes_per_elem(g) * model.elems
    pack_pass, pack_launch = model.pack_only, "direct launches"
    try:                                           # one pass = one graph launch (every K1 launch of the conversion captured once)
        model.pack_only()
        torch.cuda.synchronize(dev)
        side = torch.cuda.Stream(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            model.pack_only()
        graph.replay()
        torch.cuda.synchronize(dev)
        pack_pass, pack_launch = graph.replay, "CUDA graph replay"
    except Exception as e:                         # capture is an optimisation of the launch path only
        pack_launch = f"direct launches (graph capture failed: {str(e)[:80]})"
        torch.cuda.synchronize(dev)
    # burst: the pass alone at the clocks a kernel timed in isolation gets.  The legs above leave the GPU in its
    # power-capped state (~1.3 GHz, a moving power average); K1 in column-slab mode is bound by the SM's FMA / issue
    # rate, i.e. by the clock, so the GPU idles 1.5 s and then gets ~0.1 s of this same pass to ramp up first.  The
    # sustained figure below is the pass back to back for >= 1.2 s, until the power cap bites again (~1.75 GHz).
    burst, burst_mine = [], []
    barrier()
    time.sleep(1.5)                                # let the power-cap window of the tensor-core legs run out
    timed(pack_pass, 1, 24)
    for _ in range(5):
        ms, mine_ms = timed(pack_pass, 4, 0)
        burst.append(ms)
        burst_mine.append(mine_ms)
    ms_burst = min(burst)
    reps = max(3, int(1200.0 / max(ms_burst, 1e-3)))
    ms_sus, ms_sus_mine = timed(pack_pass, reps, 0)
    pack = {"bound": "hbm", "kernel": "group_quant_tma_cs (K1 column-slab mode, one launch per %d searched linears) + group_quant_tma "
                                      "(K1 flat mode, one launch over the arena of the other tensors)" % model.FINAL_CHUNK,
            "achieved": alg_bytes / (ms_sus * 1e-3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
            "frac": alg_bytes / (ms_sus * 1e-3) / 1e9 / peaks["hbm"], "peak_source": peaks["source"],
            "sustained": {"ms_per_pass": ms_sus, "passes": reps, "gbs_of_bf16": payload_bytes / (ms_sus * 1e-3) / 1e9},
            "burst": {"ms_per_pass": ms_burst, "achieved": alg_bytes / (ms_burst * 1e-3) / 1e9,
                      "frac": alg_bytes / (ms_burst * 1e-3) / 1e9 / peaks["hbm"],
                      "gbs_of_bf16": payload_bytes / (ms_burst * 1e-3) / 1e9},
            "algorithmic_bytes_per_pass": alg_bytes, "traffic": None, "share_of_step": ms_burst / ms_step_mine,
            "launch": pack_launch, "kernels_per_pass": model.k1_launches, "tensors_per_pass": len(model.items)}
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tpath):                      # per-launch DRAM bytes from the committed ncu --set full captures
        with open(tpath) as f:
            tj = json.load(f)
        if roofline is not None:
            t = tj.get("search_fused_kernel", {})
            roofline["traffic"] = t.get("dram_bytes_per_launch")
            roofline["traffic_detail"] = t
        t = tj.get("group_quant_tma", {})
        pack["traffic"] = t.get("dram_bytes_per_launch")
        pack["traffic_detail"] = t
    if roofline is None:                           # --no-search: the pack kernel is the dominant one
        roofline = {k: pack[k] for k in ("bound", "kernel", "achieved", "peak", "unit", "frac", "traffic", "peak_source")}

    per_rank = None
    if world > 1:
        rows = [None] * world
        dist.all_gather_object(rows, {"rank": rank, "tensors": len(shard), "searched": len(model.searched),
                                      "bf16_bytes": payload_bytes, "step_ms": round(ms_step_mine, 3),
                                      "pack_burst_ms": round(min(burst_mine), 4), "pack_sustained_ms": round(ms_sus_mine, 4),
                                      "pack_frac_sustained": round(alg_bytes / (ms_sus_mine * 1e-3) / 1e9 / peaks["hbm"], 4),
                                      "search_tflops": round(model.search_flops / (ms_scores_mine * 1e-3) / 1e12, 1) if model.searched else None})
        per_rank = rows

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
        "vs_baseline": None, "dtype": "bf16 (tcgen05 bf16 x bf16 -> f32 scores; f32 quantizer arithmetic)", "data": "synthetic",
        "config": {"workload": f"{workload}-shaped full AWQ convert: {n_grid}-point alpha search (T={T}) + int4 g{g} "
                               f"{'symmetric' if sym else 'asymmetric'} quantize + pack" if search else
                               f"{workload}-shaped plain int4 g{g} quantize + pack (no search)",
                   "tensors": len(specs), "params": total_bytes // 2, "bf16_GB": total_bytes / 1e9,
                   "searched_linears_this_rank": len(model.searched),
                   "l2": "inputs larger than L2 (%.1f GB of weights per rank per step)" % (payload_bytes / 1e9),
                   "parallelism": f"tensor-sharded x{world} (LPT), no data-path collective",
                   "timed_region_s": timed_region_s, "numa_node_rank0": numa_node, "host_numa_nodes": numa_nodes(),
                   "host_cores": os.cpu_count(), "per_rank": per_rank},
        "s_per_model": ms_step * 1e-3, "roofline": roofline, "pack": pack,
        "gpu_launches": model.launches * args.steps,
    }

    # ---- e2e leg: the public call, ordinary host tensors in, packed host results out --------------------------
    if not args.no_e2e:
        line["e2e"] = e2e_leg(torch, dist, AWQQuantizer, model, dev, world, local, g, sym, n_grid, total_bytes, allmax, barrier)
    line["clocks"] = clocks_step

    # ---- CPU baseline: the reference on this box's host cores (rank 0, N=1 only) ------------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(torch, model, g, sym, args.cpu_seconds)

    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)
    return 0


def e2e_leg(torch, dist, AWQQuantizer, model, dev, world, local, g, sym, n_grid, total_bytes, allmax, barrier):
    """AWQQuantizer.quantize_model(dict of pageable host tensors, activations=..., pack=True): the drop-in call.
    Every step copies the step's weights and activations host -> device and the packed results device -> host."""
    import psutil
    need = 2 * model.elems * 1.6 + (4 << 30)
    avail = psutil.virtual_memory().available / max(1, world)
    items = model.items
    note = None
    if need > avail:                                     # bounded host memory: a leading slice of the shard
        frac = max(0.05, avail / need * 0.8)
        keep, acc = [], 0
        for it in items:
            keep.append(it)
            acc += it[1] * it[2]
            if acc >= frac * model.elems:
                break
        items, note = keep, f"host memory bound: first {len(keep)} of {len(model.items)} tensors of the shard"
    host_w = {n: model.w[n].cpu() for n, *_ in items}                      # ordinary pageable tensors, as a loader returns
    host_x = {key: x.cpu().pin_memory() for key, x in model.x.items()}
    acts = {n: host_x[(ck, K)] for n, _, K, ck in items if ck is not None}
    elems = sum(C * K for _, C, K, _ in items)
    qz = AWQQuantizer(bits=4, group_size=g, symmetric=sym, device=f"cuda:{local}", logger_level="ERROR", n_grid=n_grid)
    times = []
    res = None
    for it in range(3):
        del res
        barrier()
        t0 = time.perf_counter()
        res = qz.quantize_model(host_w, activations=acts or None, pack=True)
        torch.cuda.synchronize(dev)
        times.append(time.perf_counter() - t0)
        assert len(res) == len(host_w), (len(res), len(host_w))
    dt = allmax(min(times[1:]))
    first = allmax(times[0])
    stats = dict(getattr(qz, "last_stream_stats", None) or {})
    pinned = None
    if world == 1 and 2 * elems <= (24 << 30):          # sub-record: the same call on PINNED input tensors (what a
        try:                                            # loader that reads into page-locked memory would hand over)
            from awq_quantizer import _native as N
            t0 = time.perf_counter()
            # one page-locked block for the whole model from awqk_host_alloc_pinned (huge pages, page-locked in place:
            # ~1 s for 16 GB; torch's pin_memory() = cudaHostAlloc takes 13 s), tensors copied in by the native copy pool
            offs, total = {}, 0
            for n, t in host_w.items():
                offs[n] = total
                total += -(-t.numel() * t.element_size() // 256) * 256
            block = N.pinned_take(total)
            pin_w = {}
            for n, t in host_w.items():
                v = block[offs[n]:offs[n] + t.numel() * t.element_size()].view(t.dtype).view(t.shape)
                N.host_copy(v, t)
                pin_w[n] = v
            pin_s = time.perf_counter() - t0
            tp = []
            for it in range(3):
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                r2 = qz.quantize_model(pin_w, activations=acts or None, pack=True)
                torch.cuda.synchronize(dev)
                tp.append(time.perf_counter() - t0)
                del r2
            pinned = {"s_per_model": min(tp[1:]), "value": 2 * elems / min(tp[1:]) / 1e9, "unit": UNIT,
                      "pinning_the_inputs_s": pin_s,
                      "note": "no staging copy: the upload DMA reads the caller's tensors (page-locked block from "
                              "awqk_host_alloc_pinned + one copy of the model into it)"}
            del pin_w
            N.pinned_give_back(block)
            del block
        except Exception as e:
            pinned = {"error": str(e)[:120]}
    d2h = sum(v.numel() * v.element_size() for r in res.values() for v in r.values() if hasattr(v, "numel") and v.dim() > 0)
    h2d = 2 * elems + sum(x.numel() * 2 for x in {id(a): a for a in acts.values()}.values())
    done_bytes = 2 * elems
    if world > 1:
        t = torch.tensor([done_bytes, h2d, d2h], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        done_bytes, h2d, d2h = (float(v) for v in t)
    return {"value": done_bytes / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "s_per_model": dt * (total_bytes / done_bytes), "s_per_step_measured": dt, "first_call_s": first,
            "steps": len(times) - 1, "api": "AWQQuantizer.quantize_model(dict of pageable host tensors, activations=..., pack=True)",
            "host_side_seconds_rank0": {k: (round(v, 4) if isinstance(v, float) else v) for k, v in stats.items()},
            "from_pinned_inputs": pinned,
            "coverage": note or "the whole shard of every rank"}


def cpu_baseline(torch, model, g, sym, seconds):
    ref = CpuReference(g, sym)
    name, C, K, _ = max(model.searched or model.items, key=lambda it: it[1] * it[2])
    w = model.w[name].cpu()
    rows = ref.rows_for(w, seconds)
    sample = w[:rows].contiguous()
    t0 = time.perf_counter()
    ref.run(sample)
    secs = time.perf_counter() - t0
    cb = {"value": sample.numel() * 2 / secs / 1e9, "unit": UNIT, "cores": torch.get_num_threads(), "kind": ref.kind,
          "sample": f"first {rows} rows of {name} ({C}x{K}) bf16, {secs:.1f} s; {ref.what}; group quantization only "
                    f"(the reference has no scale search)"}
    try:                                               # the C restatement on all cores, and the oracle's search, for scale
        from oracle import c_oracle as CO
        big = model.w[name].cpu()
        t0 = time.perf_counter()
        CO.group_quant(big, 4, g, sym, threads=os.cpu_count())
        cb["c_restatement_all_cores"] = {"value": big.numel() * 2 / (time.perf_counter() - t0) / 1e9, "unit": UNIT,
                                         "cores": os.cpu_count(), "sample": f"{name} whole tensor"}
    except Exception as e:
        cb["c_restatement_all_cores"] = {"error": str(e)[:100]}
    return cb


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

"""Activation-aware per-input-channel scale search (the "AWQ" of the north star).

The reference only cites the AWQ paper (awq.py:1-7); it has no activations, no alpha grid and no
GEMM.  This module adds them as an opt-in:
``AWQQuantizer.quantize(W, activations=X)``.  Definition = oracle/awq_oracle.py::search_scales:

    m      = mean_t |X[t, :]|                                  (fp64 accumulation)
    s_i    = clamp(m ** (i / n_grid), 1e-4) / sqrt(max * min)   i = 0 .. n_grid-1
    dW_i   = W - dequant(group_quant(W * s_i)) / s_i            the reference's group quantizer, fp32
    err_i  = mean_{t,c} (X . dW_i^T)^2                          tcgen05 bf16 GEMM, fp32 accumulate
    best   = argmin err (ties -> smallest i);  result = group_quant(W * s_best) (+ pack)

Everything runs in the sm_100a kernels of csrc/awqk_search.cu; the only host work is the argmin
over n_grid doubles.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .. import _native as N

_DW_BUDGET_BYTES = 8 << 30    # delta workspace per chunk of the alpha grid


def _check(w: torch.Tensor, x: torch.Tensor, group_size: int):
    if w.dim() != 2:
        raise ValueError(f"activation-aware search needs a 2-D weight [out, in], got {tuple(w.shape)}")
    if x.dim() != 2 or x.shape[1] != w.shape[1]:
        raise ValueError(f"activations must be [tokens, {w.shape[1]}], got {tuple(x.shape)}")
    if group_size not in (32, 64, 128) or w.shape[1] % group_size != 0:
        raise ValueError("activation-aware search needs group_size in {32, 64, 128} dividing the input dim")
    if w.shape[1] % 8 != 0:
        raise ValueError("input dim must be a multiple of 8")


def search_device(w: torch.Tensor, x: torch.Tensor, *, bits: int, group_size: int, symmetric: bool,
                  n_grid: int = 20, s_grid: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """Device-resident search.  ``w`` [C, K] (bf16/fp16/fp32) and ``x`` [T, K] are CUDA tensors.
    Returns CUDA tensors: 's_grid' fp32 [n_grid, K], 'err_sum' fp64 [n_grid] (sum, not mean),
    'act_colsum' fp64 [K].  ``s_grid`` may be injected (tests: "given equal scales")."""
    _check(w, x, group_size)
    L = N.lib()
    dev = w.device
    st = N.stream_ptr(dev)
    C, K = w.shape
    T = x.shape[0]
    colsum = torch.zeros(K, dtype=torch.float64, device=dev)
    N.check(L.awqk_abs_colsum(N.ptr(x), N.dtype_code(x.dtype), T, K, N.ptr(colsum), st), "awqk_abs_colsum")
    if s_grid is None:
        s_grid = torch.empty((n_grid, K), dtype=torch.float32, device=dev)
        ws = torch.empty(2 * n_grid, dtype=torch.float32, device=dev)
        N.check(L.awqk_alpha_grid(N.ptr(colsum), T, K, n_grid, N.ptr(s_grid), N.ptr(ws), st), "awqk_alpha_grid")
    else:
        s_grid = s_grid.to(device=dev, dtype=torch.float32).contiguous()
        n_grid = s_grid.shape[0]
    xb = x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16)
    xb = xb.contiguous()
    err = torch.zeros(n_grid, dtype=torch.float64, device=dev)
    per_alpha = C * K * 2
    chunk = max(1, min(n_grid, _DW_BUDGET_BYTES // per_alpha))
    dw = torch.empty((chunk, C, K), dtype=torch.bfloat16, device=dev)
    rws = torch.empty((chunk, K), dtype=torch.float32, device=dev)
    for a0 in range(0, n_grid, chunk):
        n_s = min(chunk, n_grid - a0)
        N.check(L.awqk_fakequant_delta(N.ptr(w), N.dtype_code(w.dtype), C, K, group_size, bits, int(symmetric),
                                       s_grid[a0].data_ptr(), n_s, N.ptr(dw), N.ptr(rws), st), "awqk_fakequant_delta")
        N.check(L.awqk_sqerr_gemm(N.ptr(xb), N.ptr(dw), T, C, K, n_s, err[a0].data_ptr() if a0 else N.ptr(err), st),
                "awqk_sqerr_gemm")
    return {"s_grid": s_grid, "err_sum": err, "act_colsum": colsum, "dw_last": dw}


def quantize_with_search(qz, tensor: torch.Tensor, activations: torch.Tensor, dev: torch.device, *,
                         pack: bool = False) -> Dict[str, torch.Tensor]:
    """``AWQQuantizer.quantize(tensor, activations=...)``: search, then the final group quantization
    of fp32 (W * s_best) with K1 (col_scale path, fp32 arithmetic)."""
    w = tensor.to(dev, non_blocking=True).contiguous()
    x = activations.to(dev, non_blocking=True).contiguous()
    r = search_device(w, x, bits=qz.bits, group_size=qz.group_size, symmetric=qz.symmetric, n_grid=qz.n_grid)
    T, C = x.shape[0], w.shape[0]
    err = (r["err_sum"] / float(T * C)).cpu()
    best = int(torch.argmin(err))                      # first minimum -> ties go to the smallest alpha
    s_best = r["s_grid"][best].contiguous()
    out = qz._quantize_device(w, pack=pack, col_scale=s_best, arith="fp32")
    out = qz._to_host(out, dev)
    result = {
        "tensor_q": out["tensor_q"],
        "scales": out["scales"],
        "zero_points": out["zero_points"],
        "bits": torch.tensor(qz.bits, dtype=torch.int32),
        "group_size": torch.tensor(qz.group_size, dtype=torch.int32),
        "symmetric": torch.tensor(qz.symmetric, dtype=torch.bool),
        "awq_scale": s_best.cpu(),
        "alpha": torch.tensor(best / qz.n_grid, dtype=torch.float32),
        "best_idx": torch.tensor(best, dtype=torch.int32),
        "search_err": err,
    }
    if pack:
        result["qweight"] = out["qweight"]
        result["qzeros"] = out["qzeros"]
    return result


class SearchPipeline:
    """Runs the search for many linears with the bandwidth-bound prologue (column statistic, alpha
    grid, fake-quant deltas) of tensor i+1 overlapped with the tensor-core GEMM of tensor i: two CUDA
    streams, two delta workspaces, events only -- no host synchronisation until ``results()``.
    The scale grid is cached per activation tensor (q/k/v or gate/up share theirs)."""

    def __init__(self, dev: torch.device, *, bits: int, group_size: int, symmetric: bool, n_grid: int = 20):
        self.dev, self.bits, self.g, self.sym, self.n_grid = dev, bits, group_size, symmetric, n_grid
        self.s_prep = torch.cuda.Stream(dev)
        self.s_gemm = torch.cuda.Stream(dev)
        self.bufs = [None, None]
        self.buf_free = [None, None]           # event: GEMM that last read this buffer has finished
        self.grid_cache = {}
        self.pending = []
        self.i = 0
        self._err_blocks, self._err_used = [], 0
        self._events, self._ev_i = [], 0
        self._last_cur_sync = None

    def _grid(self, x: torch.Tensor):
        key = (x.data_ptr(), tuple(x.shape))
        if key not in self.grid_cache:
            L = N.lib()
            T, K = x.shape
            st = self.s_prep.cuda_stream
            with torch.cuda.stream(self.s_prep):
                colsum = torch.zeros(K, dtype=torch.float64, device=self.dev)
                s_grid = torch.empty((self.n_grid, K), dtype=torch.float32, device=self.dev)
                ws = torch.empty(2 * self.n_grid, dtype=torch.float32, device=self.dev)
                rws = torch.empty((self.n_grid, K), dtype=torch.float32, device=self.dev)
                xb = (x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16)).contiguous()
            N.check(L.awqk_abs_colsum(N.ptr(x), N.dtype_code(x.dtype), T, K, N.ptr(colsum), st), "awqk_abs_colsum")
            N.check(L.awqk_alpha_grid(N.ptr(colsum), T, K, self.n_grid, N.ptr(s_grid), N.ptr(ws), st), "awqk_alpha_grid")
            self.grid_cache[key] = (s_grid, xb, rws, x)
        return self.grid_cache[key]

    def _err_row(self):
        """rows of pre-zeroed fp64 blocks (no per-tensor allocation / memset launch); returns
        (row tensor, block index, row index)"""
        if not self._err_blocks or self._err_used == self._err_blocks[-1].shape[0]:
            with torch.cuda.stream(self.s_prep):
                self._err_blocks.append(torch.zeros((256, self.n_grid), dtype=torch.float64, device=self.dev))
            self._err_used = 0
        row = self._err_blocks[-1][self._err_used]
        where = (len(self._err_blocks) - 1, self._err_used)
        self._err_used += 1
        return row, where

    def submit(self, name: str, w: torch.Tensor, x: torch.Tensor) -> None:
        _check(w, x, self.g)
        L = N.lib()
        C, K = w.shape
        T = x.shape[0]
        b = self.i & 1
        self.i += 1
        need = self.n_grid * C * K
        if self.i == 1 or self._last_cur_sync is not w:      # inputs were produced on the caller's stream
            self.s_prep.wait_stream(torch.cuda.current_stream(self.dev))
        self._last_cur_sync = w
        if self.buf_free[b] is not None:
            self.s_prep.wait_event(self.buf_free[b])
        if self.bufs[b] is None or self.bufs[b].numel() < need:
            with torch.cuda.stream(self.s_prep):
                self.bufs[b] = torch.empty(need, dtype=torch.bfloat16, device=self.dev)
        s_grid, xb, rws, _ = self._grid(x)
        err, where = self._err_row()
        sp, sg = self.s_prep.cuda_stream, self.s_gemm.cuda_stream
        N.check(L.awqk_fakequant_delta(w.data_ptr(), N.dtype_code(w.dtype), C, K, self.g, self.bits, int(self.sym),
                                       s_grid.data_ptr(), self.n_grid, self.bufs[b].data_ptr(), rws.data_ptr(), sp),
                "awqk_fakequant_delta")
        ready = self._event()
        ready.record(self.s_prep)
        self.s_gemm.wait_event(ready)
        N.check(L.awqk_sqerr_gemm(xb.data_ptr(), self.bufs[b].data_ptr(), T, C, K, self.n_grid, err.data_ptr(), sg),
                "awqk_sqerr_gemm")
        done = self._event()
        done.record(self.s_gemm)
        self.buf_free[b] = done
        self.pending.append((name, where, s_grid, float(T * C), w, x))

    def _event(self):
        if self._ev_i == len(self._events):
            self._events.append(torch.cuda.Event())
        e = self._events[self._ev_i]
        self._ev_i += 1
        return e

    def finish(self):
        """joins both streams into the current one; returns [(name, err_mean fp64[n_grid] (device),
        best_idx (device int64 0-d), s_best (device fp32 [K]))] without a host sync.  The argmin and the
        gather of the winning scale vectors are batched: one launch per error block / activation tensor."""
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_stream(self.s_prep)
        cur.wait_stream(self.s_gemm)
        if not self.pending:
            return []
        n = len(self.pending)
        denom = torch.tensor([p[3] for p in self.pending], dtype=torch.float64, device=self.dev)
        rows = []
        for bi, blk in enumerate(self._err_blocks):
            cnt = sum(1 for p in self.pending if p[1][0] == bi)
            rows.append(blk[:cnt])
        means = torch.cat(rows) / denom[:, None]                  # pending order == allocation order
        best = torch.argmin(means, dim=1)                         # first minimum -> smallest alpha on ties
        s_best = [None] * n
        by_grid = {}
        for i, p in enumerate(self.pending):
            by_grid.setdefault(id(p[2]), (p[2], []))[1].append(i)
        for s_grid, idxs in by_grid.values():
            sel = s_grid.index_select(0, best)                    # [n, K] rows for every tensor: no index upload
            for i in idxs:
                s_best[i] = sel[i]
        out = [(p[0], means[i], best[i], s_best[i]) for i, p in enumerate(self.pending)]
        self.pending = []
        self._ev_i = 0                       # events are reusable once both streams were joined
        self._err_blocks, self._err_used = [], 0
        return out


# model-level, streamed form (uploader thread / search / result ring): quantization/stream.py
from .stream import quantize_model_with_search  # noqa: E402,F401  (re-exported: the public entry of this module)


def bench_leg(args, dev, world: int, rank: int, tf_peak: float, peak_kind: str):
    """bench.py's search leg: every linear of this rank's share of the workload, synthetic
    activations X[T, K] = N(0,1) * exp(N(0,1)) per channel, n_grid = 20.  Device-resident, CUDA-event
    timed, max over ranks.  Reports s/model and the GEMM kernel against the bf16 tensor peak."""
    import torch.distributed as dist
    from .. import model_shapes as M
    L = N.lib()
    g, n_grid, T = args.group_size, 20, args.search_tokens
    specs = M.workload(args.workload)
    pool = [(f"r{r}/{name}", shape) for r in range(world) for name, shape, ck in specs if ck is not None]
    bins = M.partition_lpt([(n, M.numel(s) * T) for n, s in pool], world)     # cost ~ C*K*T
    mine = [(n, s) for n, s in pool if n in set(bins[rank])]
    distinct = sorted({tuple(s) for _, s in mine})
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    xs = {}
    for _, (C, K) in [(None, s) for s in distinct]:
        if K not in xs:
            gain = torch.exp(torch.randn(K, generator=gen, device=dev))
            xs[K] = (torch.randn((T, K), generator=gen, device=dev) * gain).to(torch.bfloat16)
    ws = {s: (torch.randn(s, generator=gen, device=dev) * 0.02).to(torch.bfloat16) for s in distinct}
    counts = {s: sum(1 for _, t in mine if tuple(t) == s) for s in distinct}

    def one(shape):
        return search_device(ws[shape], xs[shape[1]], bits=4, group_size=g, symmetric=args.symmetric, n_grid=n_grid)

    pipe = SearchPipeline(dev, bits=4, group_size=g, symmetric=args.symmetric, n_grid=n_grid)

    def whole_model():
        for s in distinct:
            for j in range(counts[s]):
                pipe.submit(f"{s}/{j}", ws[s], xs[s[1]])
        return pipe.finish()

    whole_model()           # warm-up (also allocates the delta workspaces in torch's caching allocator)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0.record()
    res = whole_model()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    flops = sum(2.0 * T * s[0] * s[1] * n_grid * counts[s] for s in distinct)
    if world > 1:
        t = torch.tensor([ms, flops], device=dev, dtype=torch.float64)
        dist.all_reduce(t[:1], op=dist.ReduceOp.MAX)
        dist.all_reduce(t[1:], op=dist.ReduceOp.SUM)
        ms, flops = float(t[0]), float(t[1])
    # the GEMM kernel alone on the largest shape (roofline of the dominant kernel of this leg)
    big = max(distinct, key=lambda s: s[0] * s[1])
    C, K = big
    r = one(big)
    dw, xb = r["dw_last"], xs[K]
    n_s = dw.shape[0]
    err = torch.zeros(n_s, dtype=torch.float64, device=dev)
    st = N.stream_ptr(dev)
    for _ in range(2):
        N.check(L.awqk_sqerr_gemm(N.ptr(xb), N.ptr(dw), T, C, K, n_s, N.ptr(err), st))
    torch.cuda.synchronize(dev)
    reps = 5
    e0.record()
    for _ in range(reps):
        N.check(L.awqk_sqerr_gemm(N.ptr(xb), N.ptr(dw), T, C, K, n_s, N.ptr(err), st))
    e1.record()
    torch.cuda.synchronize(dev)
    gemm_ms = e0.elapsed_time(e1) / reps
    gemm_tf = 2.0 * T * C * K * n_s / (gemm_ms * 1e-3) / 1e12
    return {
        "s_per_model": ms * 1e-3, "tokens": T, "n_grid": n_grid, "linears_per_rank": len(mine),
        "tflops_executed": flops / (ms * 1e-3) / 1e12,
        "flops_executed": flops, "flops_survey_formula": flops * (n_grid + 1) / n_grid,
        "roofline": {"bound": "tensor", "kernel": "sqerr_gemm2_kernel (tcgen05 cta_group::2, K2)", "achieved": gemm_tf, "peak": tf_peak,
                     "unit": "TFLOP/s", "frac": gemm_tf / tf_peak, "peak_source": peak_kind + " (sustained bf16)",
                     "shape": f"T={T} C={C} K={K} n_s={n_s}", "ms_per_launch": gemm_ms, "traffic": None},
    }

"""Activation-aware per-input-channel scale search (the "AWQ" of the north star).

The reference only cites the AWQ paper (awq.py:1-7); it has no activations, no alpha grid and no
GEMM.  This module adds them as an opt-in:
``AWQQuantizer.quantize(W, activations=X)``.  Definition = oracle/awq_oracle.py::search_scales:

    m      = mean_t |X[t, :]|                                  (fp64 accumulation)
    s_i    = clamp(m ** (i / n_grid), 1e-4) / sqrt(max * min)   i = 0 .. n_grid-1
    dW_i   = W - dequant(group_quant(W * s_i)) / s_i            the reference's group quantizer, fp32
    err_i  = mean_{t,c} (X . dW_i^T)^2                          tcgen05 bf16 GEMM, fp32 accumulate
    best   = argmin err (ties -> smallest i);  result = group_quant(W * s_best) (+ pack)

Everything -- scores, argmin, winning scale vector, final column-scaled quantization -- is ONE call of the C ABI
(awqk_scale_search, include/awqk.h); this module only allocates buffers and caches the scale grid per activation
tensor.  Nothing here synchronises with the host.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .. import _native as N


def _check(w: torch.Tensor, x: torch.Tensor, group_size: int):
    if w.dim() != 2:
        raise ValueError(f"activation-aware search needs a 2-D weight [out, in], got {tuple(w.shape)}")
    if x.dim() != 2 or x.shape[1] != w.shape[1]:
        raise ValueError(f"activations must be [tokens, {w.shape[1]}], got {tuple(x.shape)}")
    if group_size not in (32, 64, 128) or w.shape[1] % group_size != 0:
        raise ValueError("activation-aware search needs group_size in {32, 64, 128} dividing the input dim")
    if w.shape[1] % 64 != 0:
        raise ValueError("input dim must be a multiple of 64")


def activation_grid(x: torch.Tensor, n_grid: int, st: int):
    """column statistic and alpha grid of one activation tensor on stream ``st``: (colsum fp64 [K], s_grid fp32
    [n_grid, K], x as contiguous bf16)"""
    L = N.lib()
    dev = x.device
    T, K = x.shape
    colsum = torch.zeros(K, dtype=torch.float64, device=dev)
    s_grid = torch.empty((n_grid, K), dtype=torch.float32, device=dev)
    ws = torch.empty(2 * n_grid, dtype=torch.float32, device=dev)
    N.check(L.awqk_abs_colsum(N.ptr(x), N.dtype_code(x.dtype), T, K, N.ptr(colsum), st), "awqk_abs_colsum")
    N.check(L.awqk_alpha_grid(N.ptr(colsum), T, K, n_grid, N.ptr(s_grid), N.ptr(ws), st), "awqk_alpha_grid")
    xb = (x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16)).contiguous()
    return colsum, s_grid, xb


def workspace_bytes(C: int, K: int, T: int, n_grid: int):
    """(preferred, minimum) workspace of awqk_scale_search when the caller passes the scale grid"""
    import ctypes
    mn = ctypes.c_size_t(0)
    pref = N.lib().awqk_workspace_bytes(C, K, T, n_grid, 1, ctypes.byref(mn))
    return int(pref), int(mn.value)


def scale_search(w: torch.Tensor, xb: torch.Tensor, s_grid: torch.Tensor, *, bits: int, group_size: int,
                 symmetric: bool, workspace: torch.Tensor, outputs: Optional[Dict[str, torch.Tensor]] = None,
                 st: Optional[int] = None, select: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
    """ONE native call (awqk_scale_search): scores for every row of ``s_grid``, device-side argmin, winning scale
    vector and -- when ``outputs`` holds 'scales' (+ any of 'tensor_q', 'qweight', 'zero_points', 'qzeros') -- the
    final column-scaled K1 pass.  Nothing synchronises with the host."""
    dev = w.device
    C, K = w.shape
    T = xb.shape[0]
    n_grid = s_grid.shape[0]
    if select is not None:                       # caller-owned buffers ('err_mean' fp64 [n_grid], 'best_idx', 's_best')
        err, best, s_best = select["err_mean"], select["best_idx"], select["s_best"]
    else:
        err = torch.empty(n_grid, dtype=torch.float64, device=dev)
        best = torch.empty((), dtype=torch.int32, device=dev)
        s_best = torch.empty(K, dtype=torch.float32, device=dev)
    o = outputs or {}
    N.check(N.lib().awqk_scale_search(
        N.ptr(w), N.dtype_code(w.dtype), C, K, N.ptr(xb), T, N.ptr(s_grid), n_grid, group_size, bits, int(symmetric),
        N.ptr(err), N.ptr(best), N.ptr(s_best), N.ptr(o.get("tensor_q")), N.ptr(o.get("qweight")), N.ptr(o.get("scales")),
        N.ptr(o.get("zero_points")), N.ptr(o.get("qzeros")), N.ptr(workspace), workspace.numel() * workspace.element_size(),
        N.stream_ptr(dev) if st is None else st), "awqk_scale_search")
    return {"err_mean": err, "best_idx": best, "s_best": s_best}


def search_device(w: torch.Tensor, x: torch.Tensor, *, bits: int, group_size: int, symmetric: bool,
                  n_grid: int = 20, s_grid: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """Device-resident search of one tensor.  ``w`` [C, K] (bf16/fp16/fp32) and ``x`` [T, K] are CUDA tensors.
    Returns CUDA tensors: 's_grid' fp32 [n_grid, K], 'err_mean' fp64 [n_grid], 'err_sum' (= mean * T * C),
    'best_idx' int32 0-d, 's_best' fp32 [K], 'act_colsum' fp64 [K].  ``s_grid`` may be injected (tests: "given equal
    scales")."""
    _check(w, x, group_size)
    dev = w.device
    st = N.stream_ptr(dev)
    C, K = w.shape
    T = x.shape[0]
    colsum, grid, xb = activation_grid(x, n_grid, st)
    if s_grid is not None:
        grid = s_grid.to(device=dev, dtype=torch.float32).contiguous()
        n_grid = grid.shape[0]
    ws = torch.empty(workspace_bytes(C, K, T, n_grid)[0], dtype=torch.uint8, device=dev)
    r = scale_search(w.contiguous(), xb, grid, bits=bits, group_size=group_size, symmetric=symmetric, workspace=ws)
    return {"s_grid": grid, "err_mean": r["err_mean"], "err_sum": r["err_mean"] * float(T * C), "best_idx": r["best_idx"],
            "s_best": r["s_best"], "act_colsum": colsum}


def search_device_staged(w: torch.Tensor, x: torch.Tensor, s_grid: torch.Tensor, *, bits: int, group_size: int,
                         symmetric: bool) -> torch.Tensor:
    """The same scores through the stand-alone stages (awqk_fakequant_delta -> HBM -> awqk_sqerr_gemm): the
    cross-check of the fused kernel in tests / tools.  Returns err_sum fp64 [n_grid] (device)."""
    L = N.lib()
    dev = w.device
    st = N.stream_ptr(dev)
    C, K = w.shape
    T = x.shape[0]
    n_grid = s_grid.shape[0]
    xb = (x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16)).contiguous()
    err = torch.zeros(n_grid, dtype=torch.float64, device=dev)
    chunk = max(1, min(n_grid, (4 << 30) // (C * K * 2)))
    dw = torch.empty((chunk, C, K), dtype=torch.bfloat16, device=dev)
    for a0 in range(0, n_grid, chunk):
        n_s = min(chunk, n_grid - a0)
        N.check(L.awqk_fakequant_delta(N.ptr(w), N.dtype_code(w.dtype), C, K, group_size, bits, int(symmetric),
                                       s_grid[a0].data_ptr(), n_s, N.ptr(dw), st), "awqk_fakequant_delta")
        N.check(L.awqk_sqerr_gemm(N.ptr(xb), N.ptr(dw), T, C, K, n_s, err[a0:].data_ptr(), st), "awqk_sqerr_gemm")
    return err


def alloc_outputs(qz, C: int, K: int, dev: torch.device, *, pack: bool, unpacked: bool) -> Dict[str, torch.Tensor]:
    """device buffers of the final pass for one [C, K] tensor (keys as in the result dicts)"""
    G, per = K // qz.group_size, 32 // qz.bits
    o = {"scales": torch.empty((C, G), dtype=torch.float16, device=dev),
         "zero_points": torch.empty((C, G), dtype=torch.int32, device=dev)}
    if unpacked:
        o["tensor_q"] = torch.empty((C, K), dtype=torch.int32, device=dev)
    if pack:
        o["qweight"] = torch.empty((C, -(-K // per)), dtype=torch.int32, device=dev)
        o["qzeros"] = torch.empty((C, -(-G // per)), dtype=torch.int32, device=dev)
    return o


def quantize_with_search(qz, tensor: torch.Tensor, activations: torch.Tensor, dev: torch.device, *,
                         pack: bool = False) -> Dict[str, torch.Tensor]:
    """``AWQQuantizer.quantize(tensor, activations=...)``: search, then the final group quantization
    of fp32 (W * s_best) with K1 (col_scale path, fp32 arithmetic) -- one native call."""
    w = tensor.to(dev, non_blocking=True).contiguous()
    x = activations.to(dev, non_blocking=True).contiguous()
    _check(w, x, qz.group_size)
    C, K = w.shape
    T = x.shape[0]
    _, s_grid, xb = activation_grid(x, qz.n_grid, N.stream_ptr(dev))
    outs = alloc_outputs(qz, C, K, dev, pack=pack, unpacked=True)
    ws = torch.empty(workspace_bytes(C, K, T, qz.n_grid)[0], dtype=torch.uint8, device=dev)
    r = scale_search(w, xb, s_grid, bits=qz.bits, group_size=qz.group_size, symmetric=qz.symmetric, workspace=ws,
                     outputs=outs)
    host = qz._to_host({**outs, "search_err": r["err_mean"], "best_idx": r["best_idx"], "awq_scale": r["s_best"]}, dev)
    best = int(host["best_idx"])
    result = {
        "tensor_q": host["tensor_q"].reshape(tensor.shape),
        "scales": host["scales"],
        "zero_points": host["zero_points"],
        "bits": torch.tensor(qz.bits, dtype=torch.int32),
        "group_size": torch.tensor(qz.group_size, dtype=torch.int32),
        "symmetric": torch.tensor(qz.symmetric, dtype=torch.bool),
        "awq_scale": host["awq_scale"],
        "alpha": torch.tensor(best / qz.n_grid, dtype=torch.float32),
        "best_idx": host["best_idx"],
        "search_err": host["search_err"],
    }
    if pack:
        result["qweight"] = host["qweight"]
        result["qzeros"] = host["qzeros"]
    return result


class SearchPipeline:
    """Runs the search (+ final pass) for many linears on the current stream: one ``awqk_scale_search`` call per
    tensor (scores, argmin, winning scales), then ONE ``awqk_group_quant_batch`` call per ``finish()`` for the final
    column-scaled K1 pass of every tensor submitted since the last one -- a wave of linears is quantized by one
    persistent launch per 31 tensors instead of one launch per tensor.  No host synchronisation, no torch kernels
    in between.  The scale grid is cached per activation tensor (q/k/v or gate/up share theirs); one workspace
    serves every call (they are ordered by the stream)."""

    def __init__(self, dev: torch.device, *, bits: int, group_size: int, symmetric: bool, n_grid: int = 20):
        self.dev, self.bits, self.g, self.sym, self.n_grid = dev, bits, group_size, symmetric, n_grid
        self.grid_cache = {}
        self.pending = []
        self.finals = {}          # dtype code -> [(w, C, K, s_best, tensor_q, qweight, scales, zero_points, qzeros)]
        self.workspace = None

    def _grid(self, x: torch.Tensor):
        key = (x.data_ptr(), tuple(x.shape))
        if key not in self.grid_cache:
            _, s_grid, xb = activation_grid(x, self.n_grid, N.stream_ptr(self.dev))
            self.grid_cache[key] = (s_grid, xb, x)           # x kept alive: the key is its address
        return self.grid_cache[key]

    def submit(self, name: str, w: torch.Tensor, x: torch.Tensor,
               outputs: Optional[Dict[str, torch.Tensor]] = None,
               select: Optional[Dict[str, torch.Tensor]] = None) -> None:
        _check(w, x, self.g)
        C, K = w.shape
        T = x.shape[0]
        pref, _ = workspace_bytes(C, K, T, self.n_grid)
        if self.workspace is None or self.workspace.numel() < pref:
            # (a replaced workspace is returned to torch's stream-ordered allocator: safe, same stream)
            self.workspace = torch.empty(pref, dtype=torch.uint8, device=self.dev)
        s_grid, xb, _ = self._grid(x)
        r = scale_search(w, xb, s_grid, bits=self.bits, group_size=self.g, symmetric=self.sym,
                         workspace=self.workspace, outputs=None, select=select)
        if outputs is not None and outputs.get("scales") is not None:
            self.finals.setdefault(N.dtype_code(w.dtype), []).append(
                (w, C, K, r["s_best"], outputs.get("tensor_q"), outputs.get("qweight"), outputs["scales"],
                 outputs.get("zero_points"), outputs.get("qzeros")))
        self.pending.append((name, r))

    def flush_finals(self) -> None:
        """the final pass group_quant(fp32(W) * s_best) of everything submitted so far, on the current stream"""
        finals, self.finals = self.finals, {}
        for code, items in finals.items():
            N.group_quant_batch(items, code, self.g, self.bits, self.sym, N.ARITH_FP32, N.stream_ptr(self.dev))

    def drop_grid(self, x: torch.Tensor) -> None:
        """forget the cached grid of one activation tensor (its last linear has been submitted)"""
        self.grid_cache.pop((x.data_ptr(), tuple(x.shape)), None)

    def finish(self, keep_grids: bool = False):
        """[(name, err_mean fp64 [n_grid] (device), best_idx (device int32 0-d), s_best (device fp32 [K]))] in
        submission order, without a host sync"""
        self.flush_finals()
        out = [(name, r["err_mean"], r["best_idx"], r["s_best"]) for name, r in self.pending]
        self.pending = []
        if not keep_grids:
            self.grid_cache = {}
        return out


# model-level, streamed form (uploader thread / search / result ring): quantization/stream.py
from .stream import quantize_model_with_search  # noqa: E402,F401  (re-exported: the public entry of this module)

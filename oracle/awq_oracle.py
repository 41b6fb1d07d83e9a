"""CPU oracle for the AWQ-Converter quantization hot path.  TEST INFRASTRUCTURE ONLY.

This file is a restatement, in plain torch-CPU tensor arithmetic, of what the
reference (shanefitch/AWQ-Converter, ``src/awq_quantizer/quantization/awq.py``)
computes.  It is the *checker* for the CUDA kernels in
``awq-converter_b200/csrc``; nothing in the product path may import it.  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` use it.

Parity status (also in DESIGN.md):

* group quantizer, dequantizer, bf16->fp16 convert: **pinned** -- checked
  bit-for-bit against the reference itself (imported from ``/root/reference``
  in the build container) by ``tests/golden/make_golden.py``, whose frozen
  outputs live in ``tests/golden/*.npz`` and are re-checked by
  ``tests/test_oracle_golden.py`` on every run.
* nibble packing, packed zero points, activation-aware alpha search: **parity
  unpinned** -- the reference has no such code.  The
  functions below *define* them, composing the pinned group quantizer.

Reference arithmetic contract (awq.py:173-250): every operation is carried out
in the dtype of the input tensor.  On CPU torch evaluates bf16/fp16 ops in
fp32 and rounds the result of *each* op back to the storage dtype; the
expressions below keep exactly one torch op per reference op so the same
roundings happen.  ``arith='fp32'`` is simply the same function applied to
``w.float()`` (what ``ref.quantize(w.float())`` computes).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

INT32_MIN = -(2 ** 31)


# --------------------------------------------------------------------------
# a3  _calculate_qmin_qmax                                   awq.py:114-128
# --------------------------------------------------------------------------
def qrange(bits: int, symmetric: bool) -> Tuple[int, int]:
    if symmetric:
        return -(1 << (bits - 1)), (1 << (bits - 1)) - 1
    return 0, (1 << bits) - 1


def _to_int32(t: torch.Tensor) -> torch.Tensor:
    """float -> int32 as the reference's ``.to(torch.int32)`` / int32 slice
    assignment does on x86 (awq.py:363-367, 410, 412): truncation, NaN ->
    INT32_MIN (measured).  Values are already
    clamped to [qmin, qmax] unless NaN, so only the NaN case needs care."""
    nan = torch.isnan(t)
    out = torch.where(nan, torch.zeros_like(t), t).to(torch.int32)
    return torch.where(nan, torch.full_like(out, INT32_MIN), out)


# --------------------------------------------------------------------------
# a4  _compute_scale_zp_for_group                             awq.py:173-213
# --------------------------------------------------------------------------
def group_scale_zp(t_min: torch.Tensor, t_max: torch.Tensor, qmin: int, qmax: int,
                   symmetric: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """scale / zero point from a group's min and max, in the dtype of t_min.

    awq.py:196-199  symmetric: a = max(|min|,|max|); min,max = -a, a
    awq.py:202      scale = (max - min) / (qmax - qmin)
    awq.py:205      scale = clamp(scale, min=1e-10)   (1e-10 rounds to 0 in fp16)
    awq.py:207-211  zp = 0 | clamp(round(qmin - min/scale), qmin, qmax)
    """
    if symmetric:
        a = torch.maximum(t_min.abs(), t_max.abs())
        t_min, t_max = -a, a
    scale = (t_max - t_min) / (qmax - qmin)
    scale = torch.clamp(scale, min=1e-10)
    if symmetric:
        zp = torch.zeros_like(scale)
    else:
        zp = (qmin - t_min / scale).round().clamp(qmin, qmax)
    return scale, zp


# --------------------------------------------------------------------------
# a5  _quantize_tensor                                        awq.py:215-250
# --------------------------------------------------------------------------
def quantize_values(x: torch.Tensor, scale: torch.Tensor, zp: torch.Tensor,
                    qmin: int, qmax: int) -> torch.Tensor:
    """clamp(round(x / scale + zp), qmin, qmax); round is half-to-even."""
    return torch.clamp(torch.round(x / scale + zp), qmin, qmax)


def _as_rows(w: torch.Tensor) -> Tuple[torch.Tensor, int, int]:
    """awq.py:306-320: 1-D -> one channel; N-D -> dim0 channels, rest flattened."""
    if w.dim() <= 1:
        rows = w.reshape(1, -1)
    else:
        rows = w.reshape(w.shape[0], -1)
    return rows, rows.shape[0], rows.shape[1]


# --------------------------------------------------------------------------
# a6/a7/a8  _quantize_per_group + _calculate_scale_zp + quantize
#                                             awq.py:286-374, 130-171, 376-416
# --------------------------------------------------------------------------
def group_quant_raw(w: torch.Tensor, bits: int = 4, group_size: int = 128,
                    symmetric: bool = True, per_channel: bool = True):
    """Returns (q, scale, zp) all still in w's dtype (q has w's shape; scale/zp
    are [C,G], or [C] / 0-d on the numel<group_size bypass)."""
    if not w.is_floating_point():
        raise ValueError(f"Expected floating point tensor, got {w.dtype}")
    if w.numel() == 0:
        # reference: Tensor.min() / torch.stack([]) raise RuntimeError
        raise RuntimeError("cannot quantize an empty tensor")
    qmin, qmax = qrange(bits, symmetric)

    if w.numel() < group_size:                       # awq.py:297-300
        if w.dim() <= 1 or not per_channel:          # awq.py:147-149: whole tensor
            scale, zp = group_scale_zp(w.min(), w.max(), qmin, qmax, symmetric)
            q = quantize_values(w, scale, zp, qmin, qmax)
            return q, scale, zp
        flat = w.reshape(w.shape[0], -1)             # awq.py:152-171: one group per dim-0 slice
        scale, zp = group_scale_zp(flat.amin(1), flat.amax(1), qmin, qmax, symmetric)
        q = quantize_values(flat, scale[:, None], zp[:, None], qmin, qmax).reshape(w.shape)
        return q, scale, zp

    rows, C, K = _as_rows(w)
    G = math.ceil(K / group_size)                    # awq.py:323
    pad = G * group_size - K
    if pad:                                          # awq.py:337-339 zero padding joins min/max
        rows = F.pad(rows, (0, pad))
    xg = rows.reshape(C, G, group_size)
    scale, zp = group_scale_zp(xg.amin(-1), xg.amax(-1), qmin, qmax, symmetric)
    q = quantize_values(xg, scale[..., None], zp[..., None], qmin, qmax)
    q = q.reshape(C, G * group_size)[:, :K].reshape(w.shape)   # awq.py:359-368 drop the pad
    return q, scale, zp


def group_quant_vec(w: torch.Tensor, bits: int = 4, group_size: int = 128,
                    symmetric: bool = True, per_channel: bool = True,
                    arith: str = "native") -> Dict[str, torch.Tensor]:
    """Vectorised restatement of ``AWQQuantizer.quantize`` (awq.py:376-416).

    arith='native': arithmetic in w.dtype (what the reference does on that
    tensor).  arith='fp32': arithmetic in fp32 (== reference on ``w.float()``).
    """
    if not isinstance(w, torch.Tensor):
        raise ValueError(f"Expected torch.Tensor, got {type(w)}")
    if not w.is_floating_point():
        raise ValueError(f"Expected floating point tensor, got {w.dtype}")
    if arith == "fp32":
        w = w.float()
    elif arith != "native":
        raise ValueError(arith)
    q, scale, zp = group_quant_raw(w, bits, group_size, symmetric, per_channel)
    return {
        "tensor_q": _to_int32(q),                                   # awq.py:329,410
        "scales": scale.to(torch.float32).to(torch.float16),        # awq.py:327,352,411
        "zero_points": _to_int32(zp.to(torch.float32)),             # awq.py:328,353,412
        "bits": torch.tensor(bits, dtype=torch.int32),
        "group_size": torch.tensor(group_size, dtype=torch.int32),
        "symmetric": torch.tensor(symmetric, dtype=torch.bool),
    }


def group_quant_loop(w: torch.Tensor, bits: int = 4, group_size: int = 128,
                     symmetric: bool = True, per_channel: bool = True) -> Dict[str, torch.Tensor]:
    """Group-at-a-time port: one set of tiny torch ops per group, exactly the
    execution shape of the reference (awq.py:332-368: Python loop over
    channels, Python loop over groups, whole-row write-back per group).  This
    is what ``bench.py --impl reference`` times -- it is how the reference
    spends its CPU time -- and what the vectorised oracle is cross-checked
    against in tests.  Only for ``numel >= group_size`` inputs."""
    qmin, qmax = qrange(bits, symmetric)
    if w.numel() < group_size:
        return group_quant_vec(w, bits, group_size, symmetric, per_channel)
    rows, C, K = _as_rows(w)
    G = math.ceil(K / group_size)
    scales = torch.zeros((C, G))
    zps = torch.zeros((C, G))
    q_out = torch.zeros(rows.shape, dtype=torch.int32)
    pad = G * group_size - K
    for c in range(C):
        line = rows[c]
        if pad > 0:
            line = F.pad(line, (0, pad))
        blocks = line.reshape(G, group_size)
        for g in range(G):
            blk = blocks[g].contiguous()
            lo, hi = blk.min(), blk.max()
            if symmetric:
                a = max(abs(lo), abs(hi))
                lo, hi = -a, a
            s = torch.clamp((hi - lo) / (qmax - qmin), min=1e-10)
            z = torch.zeros_like(s) if symmetric else (qmin - lo / s).round().clamp(qmin, qmax)
            scales[c, g] = s
            zps[c, g] = z
            codes = torch.clamp(torch.round(blk / s + z), qmin, qmax)
            a0 = g * group_size
            a1 = min(a0 + group_size, K)
            row_view = q_out[c].reshape(-1)
            row_view[a0:a1] = codes[: a1 - a0]
            q_out[c] = row_view                      # awq.py:368 O(row) self-copy per group
    return {
        "tensor_q": q_out.reshape(w.shape),
        "scales": scales.to(torch.float16),
        "zero_points": zps.to(torch.int32),
        "bits": torch.tensor(bits, dtype=torch.int32),
        "group_size": torch.tensor(group_size, dtype=torch.int32),
        "symmetric": torch.tensor(symmetric, dtype=torch.bool),
    }


# --------------------------------------------------------------------------
# a10  dequantize / _dequantize_tensor                awq.py:459-539, 252-284
# --------------------------------------------------------------------------
def dequant_vec(qd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """(q - zp) * scale per group.  The reference multiplies an int32 group by
    a 0-d **fp16** scale, so torch type promotion makes the product fp16
    (computed in fp32, rounded once to fp16) before it is stored into the fp32
    result (awq.py:282, 496, 527-531).  Verified against the reference:
    fp32 multiplication does NOT reproduce it, fp16 does."""
    q = qd["tensor_q"]
    scales = qd["scales"]
    zps = qd["zero_points"]
    g = int(qd["group_size"].item())
    if scales.dim() != 2:
        # reference indexes scales[c, g] -> IndexError on the small-tensor layouts
        raise IndexError("too many indices for tensor of dimension %d" % scales.dim())
    rows, C, K = _as_rows(q)
    G = math.ceil(K / g)
    pad = G * g - K
    if pad:
        rows = F.pad(rows, (0, pad))
    qg = rows.reshape(C, G, g)
    diff = (qg - zps[..., None].to(torch.int32)).to(torch.float16)
    prod = (diff * scales[..., None].to(torch.float16)).to(torch.float32)
    return prod.reshape(C, G * g)[:, :K].reshape(q.shape).contiguous()


# --------------------------------------------------------------------------
# a11  convert_bf16_to_fp16                            tensor_utils.py:10-22
# --------------------------------------------------------------------------
def bf16_to_fp16(t: torch.Tensor) -> torch.Tensor:
    """bf16 -> fp16 round-to-nearest-even (overflow -> inf, subnormals kept);
    any other dtype is returned unchanged (same object)."""
    if t.dtype == torch.bfloat16:
        return t.to(torch.float16)
    return t


# --------------------------------------------------------------------------
# Packing -- PARITY UNPINNED (no reference code; SURVEY.md section 8c defines it)
# --------------------------------------------------------------------------
def pack_rows_u32(codes: torch.Tensor, qmin: int, bits: int = 4) -> torch.Tensor:
    """codes: int32 [R, N] in [qmin, qmax].  u = codes - qmin; word j of a row
    holds u[8j+i] << (4 i), i = 0..7 (bits=4) or u[4j+i] << (8 i), i = 0..3
    (bits=8); rows are padded with u = 0; a NaN code (INT32_MIN) packs as u = 0.  Returned as int32 (two's-complement
    view of the uint32 word), shape [R, ceil(N / per_word)]."""
    per = 32 // bits
    R, N = codes.shape
    u = (codes.to(torch.int64) - qmin) & ((1 << bits) - 1)
    u = torch.where(codes == INT32_MIN, torch.zeros_like(u), u)   # NaN code packs as 0 (definition)
    padn = (-N) % per
    if padn:
        u = F.pad(u, (0, padn))
    u = u.reshape(R, -1, per)
    shifts = torch.arange(per, dtype=torch.int64) * bits
    word = (u << shifts).sum(-1)
    word = torch.where(word >= 2 ** 31, word - 2 ** 32, word)
    return word.to(torch.int32)


def unpack_rows_u32(words: torch.Tensor, n: int, qmin: int, bits: int = 4) -> torch.Tensor:
    per = 32 // bits
    w = words.to(torch.int64) & 0xFFFFFFFF
    shifts = torch.arange(per, dtype=torch.int64) * bits
    u = (w[..., None] >> shifts) & ((1 << bits) - 1)
    return (u.reshape(words.shape[0], -1)[:, :n] + qmin).to(torch.int32)


def pack_result(qd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """qweight [C, ceil(K/8)] and qzeros [C, ceil(G/8)] from a quantize() dict
    whose scales are [C, G]."""
    bits = int(qd["bits"].item())
    qmin, _ = qrange(bits, bool(qd["symmetric"].item()))
    rows, C, K = _as_rows(qd["tensor_q"])
    out = dict(qd)
    out["qweight"] = pack_rows_u32(rows, qmin, bits)
    out["qzeros"] = pack_rows_u32(qd["zero_points"].reshape(C, -1), qmin, bits)
    return out


# --------------------------------------------------------------------------
# Interop layout -- PARITY UNPINNED (the reference has no exporter; this is the
# public AutoAWQ "GEMM" checkpoint layout, stated from general knowledge)
# --------------------------------------------------------------------------
AWQ_ORDER = (0, 2, 4, 6, 1, 3, 5, 7)


def _pack_awq(mat: torch.Tensor) -> torch.Tensor:
    """mat int [R, C] of 4-bit values -> int32 [R, C/8]; nibble i of word j = mat[:, 8j + AWQ_ORDER[i]]"""
    R, Cn = mat.shape
    m = mat.to(torch.int64).reshape(R, Cn // 8, 8)
    word = torch.zeros((R, Cn // 8), dtype=torch.int64)
    for i, o in enumerate(AWQ_ORDER):
        word |= (m[:, :, o] & 15) << (4 * i)
    word = torch.where(word >= 2 ** 31, word - 2 ** 32, word)
    return word.to(torch.int32)


def to_autoawq_gemm(qd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """quantize() dict of a 2-D asymmetric int4 weight [C, K] -> {'qweight' [K, C/8], 'qzeros' [G, C/8],
    'scales' [G, C]}: w[c, k] ~ (q - z) * s with q, z in [0, 15]."""
    qmin, _ = qrange(int(qd["bits"].item()), bool(qd["symmetric"].item()))
    q = qd["tensor_q"] - qmin
    z = qd["zero_points"] - qmin
    return {"qweight": _pack_awq(q.t().contiguous()), "qzeros": _pack_awq(z.t().contiguous()),
            "scales": qd["scales"].t().contiguous()}


def from_autoawq_gemm(qweight: torch.Tensor, n_out: int) -> torch.Tensor:
    """inverse of _pack_awq: int32 [R, C/8] -> int32 [R, C]"""
    w = qweight.to(torch.int64) & 0xFFFFFFFF
    out = torch.zeros((qweight.shape[0], n_out // 8, 8), dtype=torch.int64)
    for i, o in enumerate(AWQ_ORDER):
        out[:, :, o] = (w >> (4 * i)) & 15
    return out.reshape(qweight.shape[0], n_out).to(torch.int32)


# --------------------------------------------------------------------------
# Activation-aware alpha search -- PARITY UNPINNED (absent from the reference;
# public AWQ algorithm, arXiv 2306.00978, frozen here as the definition)
# --------------------------------------------------------------------------
def activation_mean(X: torch.Tensor) -> torch.Tensor:
    """m[k] = mean_t |X[t,k]|, accumulated in fp64 (exact for bf16 inputs of
    sane dynamic range, hence order independent), returned as fp32."""
    return (X.to(torch.float64).abs().sum(0) / X.shape[0]).to(torch.float32)


def alpha_scales(m: torch.Tensor, alpha: float) -> torch.Tensor:
    """s = clamp(m^alpha, 1e-4); s /= sqrt(max(s) * min(s)); fp32."""
    s = torch.clamp(m.to(torch.float32).pow(alpha), min=1e-4)
    return s / torch.sqrt(s.max() * s.min())


def fake_quant_delta(W: torch.Tensor, s: torch.Tensor, bits: int, group_size: int,
                     symmetric: bool) -> torch.Tensor:
    """dW = W - dequant(group_quant(W * s)) / s, all fp32, fp32 (un-rounded)
    scale.  Group quantizer = the pinned one, fed fp32."""
    Wf = W.to(torch.float32)
    Ws = Wf * s[None, :]
    q, scale, zp = group_quant_raw(Ws, bits, group_size, symmetric, True)
    C, K = Ws.shape
    G = scale.shape[1]
    pad = G * group_size - K
    qg = (F.pad(q, (0, pad)) if pad else q).reshape(C, G, group_size)
    deq = ((qg - zp[..., None]) * scale[..., None]).reshape(C, -1)[:, :K]
    return Wf - deq / s[None, :]


def search_error(W, X, s, bits, group_size, symmetric) -> float:
    """mean over [T, C] of (X . dW^T)^2, evaluated in fp64."""
    dW = fake_quant_delta(W, s, bits, group_size, symmetric).to(torch.float64)
    Y = X.to(torch.float64) @ dW.T
    return float((Y * Y).mean())


def search_scales(W: torch.Tensor, X: torch.Tensor, bits: int = 4, group_size: int = 128,
                  symmetric: bool = False, n_grid: int = 20,
                  s_grid: Optional[torch.Tensor] = None):
    """Full search.  alpha_i = i / n_grid, i = 0..n_grid-1; best = argmin err,
    ties -> smallest i.  ``s_grid`` ([n_grid, K] fp32) may be injected to score
    a given set of scale vectors ("given equal scales")."""
    m = activation_mean(X)
    errs, grid = [], []
    for i in range(n_grid):
        s = alpha_scales(m, i / n_grid) if s_grid is None else s_grid[i].to(torch.float32)
        grid.append(s)
        errs.append(search_error(W, X, s, bits, group_size, symmetric))
    best = min(range(n_grid), key=lambda i: (errs[i], i))
    return {"best_idx": best, "alpha": best / n_grid, "err": errs, "s_grid": torch.stack(grid),
            "s_best": grid[best], "act_mean": m}


def quantize_scaled(W: torch.Tensor, s: torch.Tensor, bits: int = 4, group_size: int = 128,
                    symmetric: bool = False) -> Dict[str, torch.Tensor]:
    """Final AWQ quantization: the pinned group quantizer on fp32 (W * s)."""
    return group_quant_vec(W.to(torch.float32) * s[None, :].to(torch.float32), bits, group_size,
                           symmetric, True, arith="native")

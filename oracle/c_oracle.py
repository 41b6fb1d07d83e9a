"""ctypes wrapper of oracle/libawq_oracle.so (the C restatement).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import torch

from . import build_oracle

_lib = None
_DT = {torch.bfloat16: 0, torch.float16: 1, torch.float32: 2}


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_oracle.build())
        i64, vp, i = C.c_int64, C.c_void_p, C.c_int
        _lib.awq_oracle_group_quant.argtypes = [vp, i, i64, i64, i, i, i, i, vp, vp, vp, i]
        _lib.awq_oracle_pack_rows.argtypes = [vp, i64, i64, i, i, vp, i]
        _lib.awq_oracle_dequant.argtypes = [vp, vp, vp, i64, i64, i, vp, i]
        _lib.awq_oracle_bf16_to_fp16.argtypes = [vp, vp, i64, i]
    return _lib


def group_quant(w: torch.Tensor, bits=4, group_size=128, symmetric=True, arith="native", threads=None):
    """[C, K]-shaped (numel >= group_size) CPU tensor -> reference-layout dict."""
    threads = threads or os.cpu_count() or 1
    w = w.contiguous()
    rows = w.reshape(1, -1) if w.dim() <= 1 else w.reshape(w.shape[0], -1)
    Cn, K = rows.shape
    G = -(-K // group_size)
    q = torch.empty(rows.shape, dtype=torch.int32)
    scales = torch.empty((Cn, G), dtype=torch.float16)
    zp = torch.empty((Cn, G), dtype=torch.int32)
    rc = lib().awq_oracle_group_quant(rows.data_ptr(), _DT[w.dtype], Cn, K, group_size, bits, int(symmetric),
                                      int(arith == "fp32"), q.data_ptr(), scales.data_ptr(), zp.data_ptr(), threads)
    assert rc == 0
    return {"tensor_q": q.reshape(w.shape), "scales": scales, "zero_points": zp,
            "bits": torch.tensor(bits, dtype=torch.int32), "group_size": torch.tensor(group_size, dtype=torch.int32),
            "symmetric": torch.tensor(symmetric, dtype=torch.bool)}


def pack_rows(codes: torch.Tensor, qmin: int, bits: int = 4, threads=None):
    threads = threads or os.cpu_count() or 1
    codes = codes.contiguous()
    R, Nn = codes.shape
    per = 32 // bits
    out = torch.empty((R, -(-Nn // per)), dtype=torch.int32)
    lib().awq_oracle_pack_rows(codes.data_ptr(), R, Nn, bits, qmin, out.data_ptr(), threads)
    return out


def dequant(qd, threads=None):
    threads = threads or os.cpu_count() or 1
    q = qd["tensor_q"].contiguous()
    rows = q.reshape(1, -1) if q.dim() <= 1 else q.reshape(q.shape[0], -1)
    out = torch.empty(rows.shape, dtype=torch.float32)
    lib().awq_oracle_dequant(rows.data_ptr(), qd["scales"].contiguous().data_ptr(),
                             qd["zero_points"].contiguous().data_ptr(), rows.shape[0], rows.shape[1],
                             int(qd["group_size"]), out.data_ptr(), threads)
    return out.reshape(q.shape)


def bf16_to_fp16(t: torch.Tensor, threads=None):
    threads = threads or os.cpu_count() or 1
    t = t.contiguous()
    out = torch.empty(t.shape, dtype=torch.float16)
    lib().awq_oracle_bf16_to_fp16(t.data_ptr(), out.data_ptr(), t.numel(), threads)
    return out

"""ctypes binding of libawqk.so (the C ABI declared in include/awqk.h).

The library is built in-tree by ``awq-converter_b200/build.py`` (``__graft_entry__.build()``).
Loading fails loudly when it is missing: there is no fallback implementation."""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libawqk.so")

BF16, FP16, FP32, FP64 = 0, 1, 2, 3
ARITH_NATIVE, ARITH_FP32 = 0, 1

_lock = threading.Lock()
_lib = None

_i64, _int, _vp = C.c_int64, C.c_int, C.c_void_p

# name -> (restype, argtypes); must list every function include/awqk.h declares
SIGNATURES = {
    "awqk_version": (_int, []),
    "awqk_error_string": (C.c_char_p, [_int]),
    "awqk_last_cuda_error": (C.c_char_p, []),
    "awqk_group_quant": (_int, [_vp, _int, _i64, _i64, _int, _int, _int, _int, _vp, _vp, _vp, _vp, _vp,
                                _vp, _vp]),
    "awqk_group_quant_path": (_int, [_int, _i64, _i64, _int, _int, _int, _vp]),
    "awqk_group_quant_batch": (_int, [_vp, _int, _int, _int, _int, _int, _int, _vp]),
    "awqk_group_quant_batch_plan": (_int, [C.POINTER(_i64), C.POINTER(_i64), _int, _int, _int, _int, C.POINTER(_i64),
                                           C.POINTER(_i64), _int]),
    "awqk_dequant": (_int, [_vp, _vp, _vp, _i64, _i64, _int, _vp, _vp]),
    "awqk_dequant_packed": (_int, [_vp, _vp, _vp, _i64, _i64, _int, _int, _int, _vp, _vp]),
    "awqk_bf16_to_fp16": (_int, [_vp, _vp, _i64, _vp]),
    "awqk_abs_colsum": (_int, [_vp, _int, _i64, _i64, _vp, _vp]),
    "awqk_alpha_grid": (_int, [_vp, _i64, _i64, _int, _vp, _vp, _vp]),
    "awqk_fakequant_delta": (_int, [_vp, _int, _i64, _i64, _int, _int, _int, _vp, _int, _vp, _vp]),
    "awqk_scale_search": (_int, [_vp, _int, _i64, _i64, _vp, _i64, _vp, _int, _int, _int, _int, _vp, _vp, _vp, _vp,
                                 _vp, _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "awqk_workspace_bytes": (C.c_size_t, [_i64, _i64, _i64, _int, _int, C.POINTER(C.c_size_t)]),
    "awqk_sqerr_gemm": (_int, [_vp, _vp, _i64, _i64, _i64, _int, _vp, _vp]),
    "awqk_export_autoawq": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _int, _vp, _vp, _vp, _vp]),
    "awqk_host_prefault": (_int, [_vp, C.c_size_t, _int]),
    "awqk_pipe_create": (_int, [_int, C.c_size_t, C.POINTER(_vp)]),
    "awqk_pipe_destroy": (None, [_vp]),
    "awqk_pipe_quant_host": (_int, [_vp, _vp, _int, _i64, _i64, _int, _int, _int, _int, _vp, _vp, _vp,
                                    _vp, _vp]),
    "awqk_pipe_quant_gather": (_int, [_vp, _int, C.POINTER(_vp), C.POINTER(_i64), _i64, _int, _int, _int, _int, _int,
                                      _vp, _vp, _vp, _vp, _vp]),
    "awqk_host_copy": (_int, [_vp, _vp, C.c_size_t, _int]),
    "awqk_host_alloc_pinned": (_int, [C.c_size_t, C.POINTER(_vp)]),
    "awqk_host_free_pinned": (_int, [_vp]),
    "awqk_pipe_sync": (_int, [_vp]),
}


class QuantItem(C.Structure):
    """awqk_quant_item of include/awqk.h"""
    _fields_ = [("w", _vp), ("C", _i64), ("K", _i64), ("col_scale", _vp), ("q_unpacked", _vp), ("q_packed", _vp),
                ("scales_f16", _vp), ("zp", _vp), ("zp_packed", _vp)]


class NativeError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """The loaded library (loaded once; thread-safe)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if "AWQK_PIPE_THREADS" not in os.environ:
                    # one process per GPU (torchrun): the ranks share the host's cores -- size the pipeline's
                    # staging-copy pool accordingly (the library reads the variable once, at first use)
                    try:
                        ranks = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
                    except ValueError:
                        ranks = 1
                    if ranks > 1:
                        os.environ["AWQK_PIPE_THREADS"] = str(max(2, min(12, (os.cpu_count() or 8) // ranks)))
                if not os.path.exists(LIB_PATH):
                    raise NativeError(
                        f"{LIB_PATH} not found: build it with `python awq-converter_b200/build.py` "
                        "(or __graft_entry__.build()).  awq_quantizer has no CPU fallback.")
                handle = C.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = handle
    return _lib


def check(rc: int, what: str = "awqk call") -> None:
    if rc == 0:
        return
    L = lib()
    msg = L.awqk_error_string(rc).decode()
    detail = L.awqk_last_cuda_error().decode() if rc == -3 else ""
    raise NativeError(f"{what} failed: {msg} ({rc})" + (f": {detail}" if detail else ""))


def dtype_code(torch_dtype) -> int:
    import torch
    table = {torch.bfloat16: BF16, torch.float16: FP16, torch.float32: FP32, torch.float64: FP64}
    if torch_dtype not in table:
        raise ValueError(f"Unsupported floating point dtype for the B200 kernels: {torch_dtype}")
    return table[torch_dtype]


def ptr(t) -> int:
    """device/host address of a tensor (None -> NULL)"""
    return None if t is None else t.data_ptr()


def stream_ptr(device) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream


def host_copy(dst, src) -> None:
    """dst.copy_(src) for contiguous CPU tensors of equal dtype / size through the native multi-threaded memcpy
    (torch's copy_ is single-threaded under torchrun's OMP_NUM_THREADS=1 and holds no more than one core)"""
    if dst.dtype != src.dtype or dst.numel() != src.numel() or not dst.is_contiguous() or not src.is_contiguous():
        dst.copy_(src)
        return
    check(lib().awqk_host_copy(dst.data_ptr(), src.data_ptr(), dst.numel() * dst.element_size(), 0), "awqk_host_copy")


def group_quant_batch(items, dtype_code_: int, group_size: int, bits: int, symmetric: bool, arith: int, stream) -> None:
    """awqk_group_quant_batch over `items` = [(w, C, K, col_scale, q_unpacked, q_packed, scales, zp, zp_packed)] of
    tensors / None: the K1 pass of a whole wave of tensors in as few launches as possible."""
    if not items:
        return
    arr = (QuantItem * len(items))()
    for a, (w, Cc, K, cs, qu, qp, sc, zp, zq) in zip(arr, items):
        a.w, a.C, a.K, a.col_scale = w.data_ptr(), Cc, K, ptr(cs)
        a.q_unpacked, a.q_packed, a.scales_f16, a.zp, a.zp_packed = ptr(qu), ptr(qp), sc.data_ptr(), ptr(zp), ptr(zq)
    check(lib().awqk_group_quant_batch(C.cast(arr, _vp), len(items), dtype_code_, group_size, bits, int(symmetric), arith,
                                       stream), "awqk_group_quant_batch")


# ---- page-locked staging buffers (awqk_host_alloc_pinned) with a small process-wide cache -------------------------
# torch.empty(pin_memory=True) is cudaHostAlloc: ~0.1 s per 256 MB on these hosts.  The native allocator page-locks a
# pre-faulted huge-page mapping in place (~15 ms per 256 MB).  Buffers are handed out as uint8 tensors and come back to
# the cache when their user is done; the cache keeps at most _PIN_CACHE_BYTES and frees the oldest blocks beyond that.
_PIN_CACHE_BYTES = 3 << 30
_pin_lock = threading.Lock()
_pin_free = []            # [(nbytes, ptr, tensor)] oldest first
_pin_live = {}            # ptr -> nbytes of blocks currently handed out


def pinned_take(nbytes: int):
    """a page-locked uint8 tensor of exactly ``nbytes`` (from the cache when one of that size is free)"""
    import torch
    nbytes = max(int(nbytes), 256)
    with _pin_lock:
        for i, (nb, ptr, t) in enumerate(_pin_free):
            if nb == nbytes:
                _pin_free.pop(i)
                _pin_live[ptr] = nb
                return t
    out = _vp()
    check(lib().awqk_host_alloc_pinned(nbytes, C.byref(out)), "awqk_host_alloc_pinned")
    t = torch.frombuffer((C.c_ubyte * nbytes).from_address(out.value), dtype=torch.uint8)
    with _pin_lock:
        _pin_live[out.value] = nbytes
    return t


def pinned_give_back(t) -> None:
    """return a tensor obtained from pinned_take (all copies that use it must have finished)"""
    if t is None:
        return
    ptr = t.data_ptr()
    drop = []
    with _pin_lock:
        nb = _pin_live.pop(ptr, None)
        if nb is None:
            return                                  # not one of ours (e.g. torch's own pinned allocation)
        if nb > _PIN_CACHE_BYTES // 2:              # too big to keep: it would push every staging slot out of the cache
            drop.append(ptr)
            nb = None
    if nb is None:
        for p in drop:
            lib().awqk_host_free_pinned(p)
        return
    with _pin_lock:
        _pin_free.append((nb, ptr, t))
        total = sum(b for b, _, _ in _pin_free)
        while total > _PIN_CACHE_BYTES and len(_pin_free) > 1:
            b, p, _t = _pin_free.pop(0)
            total -= b
            drop.append(p)
    for p in drop:
        lib().awqk_host_free_pinned(p)

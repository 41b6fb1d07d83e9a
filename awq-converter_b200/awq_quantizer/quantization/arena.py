"""Whole-model quantize+pack through ONE flat arena per dtype.

The reference walks the model one tensor at a time (main.py:353-392 -> awq.py:376): move to the
device, loop over groups, copy back.  Every tensor whose rows are a whole number of groups is, for
K1, just a flat run of independent groups -- so a model is too, once its tensors sit back to back in
one buffer with tile-aligned starts.  `HostArena` is that buffer (pinned host memory, 8192-element
aligned slots = K1's CTA tile); `quantize_arena` streams it through `awqk_pipe_quant_host`
(chunked H2D -> K1 -> D2H on three streams) with a single C-ABI call per dtype, and hands back
per-tensor views of the packed output arenas.  No per-tensor launches, no per-tensor allocations.
"""
from __future__ import annotations

import ctypes as C
import math
import threading
from typing import Dict, List, Optional, Tuple

import torch

from .. import _native as N

TILE = 8192  # elements; K1's CTA tile and the pipeline's chunk quantum


def pipe_eligible(shape, dtype, group_size: int, bits: int) -> bool:
    """the host pipeline (awqk_pipe_quant_host) handles every tensor whose rows are whole groups"""
    if dtype not in (torch.bfloat16, torch.float16, torch.float32):
        return False
    if group_size not in (32, 64, 128):
        return False
    n = 1
    for s in shape:
        n *= s
    if n < group_size or n == 0:
        return False
    rows = 1 if len(shape) <= 1 else shape[0]
    return (n // rows) % group_size == 0


def arena_eligible(shape, dtype, group_size: int, bits: int) -> bool:
    """flat layout + flat zero-point packing: rows are whole groups AND whole packed zero words"""
    if not pipe_eligible(shape, dtype, group_size, bits):
        return False
    n = 1
    for s in shape:
        n *= s
    rows = 1 if len(shape) <= 1 else shape[0]
    return ((n // rows) // group_size) % (32 // bits) == 0


def short_row_len(shape, dtype, group_size: int, bits: int) -> int:
    """row length K of a tensor whose rows hold 1, 2 or 4 groups -- fewer than one packed zero word (K = 512 at
    g = 128: OPT's embed_tokens / project_in) -- else 0.  Such tensors go through the gather pipeline in their
    own class: K1 pads one zero word per row itself."""
    if not pipe_eligible(shape, dtype, group_size, bits) or arena_eligible(shape, dtype, group_size, bits):
        return 0
    n = 1
    for s in shape:
        n *= s
    rows = 1 if len(shape) <= 1 else shape[0]
    k = n // rows
    gr, per = k // group_size, 32 // bits
    return k if (gr < per and per % gr == 0 and TILE % k == 0) else 0


class HostArena:
    """Pinned host buffers (one per dtype) with a tile-aligned slot per tensor.

    ``specs``: name -> (shape, dtype).  ``views[name]`` is a tensor view into the arena: a loader can
    read file bytes straight into it (zero copy), or ``from_tensors`` copies existing tensors in."""

    def __init__(self, specs: Dict[str, Tuple[tuple, torch.dtype]], pin: bool = True, allocate: bool = True):
        """Allocates the buffers; only the tile padding behind each slot is zeroed (the slots themselves are
        about to be overwritten by the caller, and a memset of the whole model is a wasted pass over DRAM).
        ``allocate=False``: layout only (the virtual arena of awqk_pipe_quant_gather)."""
        self.specs = dict(specs)
        self.layout: Dict[torch.dtype, List[Tuple[str, int, int]]] = {}   # dtype -> [(name, offset, numel)]
        self.buffers: Dict[torch.dtype, torch.Tensor] = {}
        self.views: Dict[str, torch.Tensor] = {}
        sizes: Dict[torch.dtype, int] = {}
        for name, (shape, dtype) in self.specs.items():
            n = int(math.prod(shape)) if len(shape) else 1
            off = sizes.get(dtype, 0)
            self.layout.setdefault(dtype, []).append((name, off, n))
            sizes[dtype] = off + (n + TILE - 1) // TILE * TILE
        self.sizes = sizes
        if not allocate:
            return
        for dtype, total in sizes.items():
            buf = torch.empty(total, dtype=dtype, pin_memory=pin and torch.cuda.is_available())
            self.buffers[dtype] = buf
            for name, off, n in self.layout[dtype]:
                self.views[name] = buf[off:off + n].view(self.specs[name][0])
                buf[off + n:off + (n + TILE - 1) // TILE * TILE].zero_()

    @classmethod
    def from_tensors(cls, tensors: Dict[str, torch.Tensor], pin: bool = True) -> "HostArena":
        arena = cls({k: (tuple(v.shape), v.dtype) for k, v in tensors.items()}, pin=pin)
        for k, v in tensors.items():
            arena.views[k].copy_(v)
        return arena

    @classmethod
    def for_tensors(cls, tensors: Dict[str, torch.Tensor], pin: bool = True) -> "HostArena":
        """layout only -- a virtual arena: quantize_arena(..., sources=tensors) lets the native pipeline gather
        the tensors from where they are (pageable memory) through its own bounded ring of pinned buffers"""
        return cls({k: (tuple(v.shape), v.dtype) for k, v in tensors.items()}, pin=pin, allocate=False)

    def nbytes(self) -> int:
        return sum(b.numel() * b.element_size() for b in self.buffers.values())

    def payload_bytes(self) -> int:
        return sum(n * torch.empty((), dtype=dt).element_size() for dt, lay in self.layout.items() for _, _, n in lay)


class _ThreadPipes(dict):
    """the pipes of one host thread; destroyed (streams, device slots, pinned rings) when the thread goes away"""

    def __del__(self):
        try:
            L = N.lib()
            for handle in self.values():
                L.awqk_pipe_destroy(handle)
        except Exception:       # interpreter shutdown
            pass


class _PipeHandle:
    """per-thread, per-device awqk_pipe (the C object is not thread-safe by design)"""
    _tls = threading.local()

    @classmethod
    def get(cls, device_index: int, chunk_bytes: int) -> int:
        pipes = getattr(cls._tls, "pipes", None)
        if pipes is None:
            pipes = cls._tls.pipes = _ThreadPipes()
        key = (device_index, chunk_bytes)
        if key not in pipes:
            handle = C.c_void_p()
            N.check(N.lib().awqk_pipe_create(device_index, chunk_bytes, C.byref(handle)), "awqk_pipe_create")
            pipes[key] = handle
        return pipes[key]


def quantize_arena(arena: HostArena, *, bits: int, group_size: int, symmetric: bool, arith: str,
                   device: torch.device, chunk_bytes: int = 32 << 20, want_zero_points: bool = False,
                   out: Optional[dict] = None, sync: bool = True, packed: bool = True,
                   unpacked: bool = False, sources: Optional[Dict[str, torch.Tensor]] = None,
                   pin_results: bool = True, row_len: int = 0) -> Dict[str, Dict[str, torch.Tensor]]:
    """Quantize every tensor of ``arena`` through the chunked H2D -> K1 -> D2H pipeline.

    ``sources`` (name -> CPU tensor): the arena is virtual (``HostArena.for_tensors``); the native pipeline
    (awqk_pipe_quant_gather) gathers the tensors chunk by chunk through its own pinned bounce ring and drains
    the results into ordinary (pageable) host arrays -- no pinned allocation proportional to the model
    (cudaHostAlloc runs at ~2.3 GB/s, 20x slower than the pipeline itself), blocking.  With ``pin_results``
    the result arrays are pinned and written by the D2H copies directly (worth it when the allocation is re-used).
    ``row_len`` (gather mode): all tensors have rows of this many elements = 1, 2 or 4 groups (``short_row_len``);
    their ``qzeros`` are one zero-padded word per row.

    ``packed``   -> 'qweight' / 'qzeros' (+ 'scales'), ``unpacked`` -> the reference's 'tensor_q' int32 /
    'zero_points' int32 (+ 'scales').  The tensors are views of pinned host output arenas (``out`` may
    carry those arenas across calls to avoid re-allocating them)."""
    per = 32 // bits
    L = N.lib()
    pipe = _PipeHandle.get(device.index if device.index is not None else torch.cuda.current_device(), chunk_bytes)
    results: Dict[str, Dict[str, torch.Tensor]] = {}
    outs = out if out is not None else {}
    want_z = want_zero_points or unpacked
    for dtype in arena.layout:
        buf = arena.buffers.get(dtype)
        n = arena.sizes[dtype]
        pin = torch.cuda.is_available() and (sources is None or pin_results)
        if row_len and sources is None:
            raise ValueError("row_len needs the gather mode (sources=...)")
        n_zq = n // row_len if row_len else n // group_size // per
        key = (dtype, n, bits, group_size, want_z, packed, unpacked, pin, row_len)
        if key not in outs:
            outs[key] = {
                "q": torch.empty(n // per, dtype=torch.int32, pin_memory=pin) if packed else None,
                "s": torch.empty(n // group_size, dtype=torch.float16, pin_memory=pin),
                "zq": torch.empty(n_zq, dtype=torch.int32, pin_memory=pin) if packed else None,
                "z": torch.empty(n // group_size, dtype=torch.int32, pin_memory=pin) if want_z else None,
                "tq": torch.empty(n, dtype=torch.int32, pin_memory=pin) if unpacked else None,
            }
        o = outs[key]
        ar = N.ARITH_FP32 if arith == "fp32" else N.ARITH_NATIVE
        if sources is None:
            N.check(L.awqk_pipe_quant_host(pipe, buf.data_ptr(), N.dtype_code(dtype), 1, n, group_size, bits,
                                           int(symmetric), ar, N.ptr(o["tq"]), N.ptr(o["q"]), o["s"].data_ptr(),
                                           N.ptr(o["z"]), N.ptr(o["zq"])), "awqk_pipe_quant_host")
        else:
            lay = arena.layout[dtype]
            keep = [sources[name].detach().contiguous() for name, _, _ in lay]     # alive during the call
            ptrs = (C.c_void_p * len(lay))(*[t.data_ptr() for t in keep])
            nums = (C.c_int64 * len(lay))(*[numel for _, _, numel in lay])
            N.check(L.awqk_pipe_quant_gather(pipe, len(lay), ptrs, nums, row_len, N.dtype_code(dtype), group_size, bits,
                                             int(symmetric), ar, N.ptr(o["tq"]), N.ptr(o["q"]), o["s"].data_ptr(),
                                             N.ptr(o["z"]), N.ptr(o["zq"])), "awqk_pipe_quant_gather")
            del keep
        for name, off, numel in arena.layout[dtype]:
            shape = arena.specs[name][0]
            rows = 1 if len(shape) <= 1 else shape[0]
            k = numel // rows
            g = k // group_size
            r = {}
            if unpacked:
                r["tensor_q"] = o["tq"][off:off + numel].view(shape)
            r["scales"] = o["s"][off // group_size:(off + numel) // group_size].view(rows, g)
            if want_z:
                r["zero_points"] = o["z"][off // group_size:(off + numel) // group_size].view(rows, g)
            r["bits"] = torch.tensor(bits, dtype=torch.int32)
            r["group_size"] = torch.tensor(group_size, dtype=torch.int32)
            r["symmetric"] = torch.tensor(symmetric, dtype=torch.bool)
            if packed:
                r["qweight"] = o["q"][off // per:(off + numel) // per].view(rows, k // per)
                if row_len:
                    r["qzeros"] = o["zq"][off // row_len:(off + numel) // row_len].view(rows, 1)
                else:
                    r["qzeros"] = o["zq"][off // group_size // per:(off + numel) // group_size // per].view(rows, g // per)
            results[name] = r
    if sync:
        N.check(L.awqk_pipe_sync(pipe), "awqk_pipe_sync")
    return results


def quantize_rows_pipelined(t: torch.Tensor, *, bits: int, group_size: int, symmetric: bool, arith: str,
                            device: torch.device, chunk_bytes: int = 32 << 20, sync: bool = True) -> Dict[str, torch.Tensor]:
    """One host tensor whose rows are whole groups but not whole packed-zero words (e.g. K = 512 at
    g = 128): chunked by rows through the same pipeline; packed zero points are padded per row."""
    per = 32 // bits
    pin = torch.cuda.is_available()
    src = t.contiguous()
    if pin and not src.is_pinned():
        staged = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
        staged.copy_(src)
        src = staged
    rows = 1 if src.dim() <= 1 else src.shape[0]
    k = src.numel() // rows
    g = k // group_size
    out = {
        "qweight": torch.empty((rows, k // per), dtype=torch.int32, pin_memory=pin),
        "qzeros": torch.empty((rows, -(-g // per)), dtype=torch.int32, pin_memory=pin),
        "scales": torch.empty((rows, g), dtype=torch.float16, pin_memory=pin),
        "bits": torch.tensor(bits, dtype=torch.int32),
        "group_size": torch.tensor(group_size, dtype=torch.int32),
        "symmetric": torch.tensor(symmetric, dtype=torch.bool),
    }
    pipe = _PipeHandle.get(device.index if device.index is not None else torch.cuda.current_device(), chunk_bytes)
    L = N.lib()
    N.check(L.awqk_pipe_quant_host(pipe, src.data_ptr(), N.dtype_code(src.dtype), rows, k, group_size, bits,
                                   int(symmetric), N.ARITH_FP32 if arith == "fp32" else N.ARITH_NATIVE, None,
                                   out["qweight"].data_ptr(), out["scales"].data_ptr(), None, out["qzeros"].data_ptr()),
            "awqk_pipe_quant_host")
    out["_keepalive"] = src
    if sync:
        N.check(L.awqk_pipe_sync(pipe), "awqk_pipe_sync")
    return out


def sync_pipe(device: torch.device, chunk_bytes: int = 32 << 20) -> None:
    pipe = _PipeHandle.get(device.index if device.index is not None else torch.cuda.current_device(), chunk_bytes)
    N.check(N.lib().awqk_pipe_sync(pipe), "awqk_pipe_sync")

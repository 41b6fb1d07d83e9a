"""AWQQuantizer -- same class, constructor, methods and result layout as the reference
(src/awq_quantizer/quantization/awq.py:24-539), with the per-group Python loops replaced by the
sm_100a kernels behind include/awqk.h.

What is kept bit-for-bit (checked in tests/ against the reference's own outputs):
  * ``quantize(t)``   -> {'tensor_q' int32 (t.shape), 'scales' fp16 [C,G], 'zero_points' int32 [C,G],
                          'bits', 'group_size', 'symmetric'} on CPU            (awq.py:376-416)
  * arithmetic in the dtype of ``t`` (bf16 in -> bf16-rounded ops, ...)        (awq.py:192-248)
  * zero padding of ragged rows, 1-D / N-D handling, the numel < group_size bypass with [C]-shaped
    or 0-d scales                                                              (awq.py:297-339)
  * ``dequantize`` incl. its fp16 multiply and its IndexError on bypass dicts  (awq.py:459-539)
  * ``quantize_model`` log-and-skip semantics                                  (awq.py:435-457)
  * ValueError texts of parameter validation                                   (awq.py:95-112)

What is added (opt-in; defaults reproduce the reference):
  * ``arith='fp32'``            fp32 arithmetic == reference on ``t.float()``
  * ``quantize(t, pack=True)``  adds 'qweight' / 'qzeros' (8 nibbles per int32)
  * ``quantize(t, activations=X)``  activation-aware alpha search (tcgen05 kernels), see search.py

What is deliberately different: there is NO CPU execution path.  The reference silently falls back
to CPU when CUDA is missing (awq.py:76-77); here ``quantize`` raises instead.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

from .. import _native as N
from ..utils.logger import get_logger


class AWQQuantizer:
    def __init__(
        self,
        bits: int = 4,
        group_size: int = 128,
        symmetric: bool = True,
        zero_point: str = "minmax",
        percentile: float = 0.99,
        scale_method: str = "mse",
        per_channel: bool = True,
        device: Optional[str] = None,
        logger_name: str = "awq_quantizer",
        logger_level: str = "INFO",
        logger_to_file: bool = False,
        logger_file_path: Optional[str] = None,
        *,
        arith: str = "native",
        n_grid: int = 20,
        pin_results: Optional[bool] = None,
    ):
        self.bits = bits
        self.group_size = group_size
        self.symmetric = symmetric
        self.zero_point = zero_point
        self.percentile = percentile
        self.scale_method = scale_method
        self.per_channel = per_channel
        self.arith = arith
        self.n_grid = n_grid
        # quantize_model() results in page-locked host memory?  Pinning runs at ~2.3 GB/s (20x slower than the
        # pipeline), so it only pays when the allocation is re-used.  None = adaptive: the first model of this
        # quantizer gets ordinary (pageable) results drained through the pipeline's bounded pinned ring, later
        # models get pinned results straight from the D2H copies (torch's pinned allocator caches the blocks).
        # Searched tensors always go through the ring unless this is True (their drain hides under the search).
        self.pin_results = pin_results
        self._models_done = 0
        # awq.py:70-73: default device is CUDA.  (No silent CPU downgrade here.)
        self.device = "cuda" if device is None else device
        self.logger = get_logger(name=logger_name, level=logger_level, to_file=logger_to_file,
                                 file_path=logger_file_path)
        self._validate_parameters()
        self.qmin, self.qmax = self._calculate_qmin_qmax()
        self.logger.info(
            f"Initialized AWQ Quantizer with bits={bits}, group_size={group_size}, symmetric={symmetric}")
        self.logger.info(f"Quantization range: [{self.qmin}, {self.qmax}]")

    # ------------------------------------------------------------------ awq.py:95-128
    def _validate_parameters(self) -> None:
        if self.bits not in [4, 8]:
            raise ValueError(f"Unsupported bit width: {self.bits}. Supported: 4, 8.")
        if self.group_size <= 0 or not isinstance(self.group_size, int):
            raise ValueError(f"Group size must be a positive integer: {self.group_size}")
        if self.zero_point not in ["none", "minmax", "percentile"]:
            raise ValueError(f"Unsupported zero point calibration method: {self.zero_point}")
        if self.zero_point == "percentile" and (self.percentile <= 0 or self.percentile >= 1):
            raise ValueError(f"Percentile must be in range (0, 1): {self.percentile}")
        if self.scale_method not in ["minmax", "mse"]:
            raise ValueError(f"Unsupported scale calibration method: {self.scale_method}")
        if self.arith not in ("native", "fp32"):
            raise ValueError(f"Unsupported arithmetic mode: {self.arith}")

    def _calculate_qmin_qmax(self) -> Tuple[int, int]:
        if self.symmetric:
            return -(2 ** (self.bits - 1)), 2 ** (self.bits - 1) - 1
        return 0, 2 ** self.bits - 1

    # ------------------------------------------------------------------ device plumbing
    def _cuda_device(self) -> torch.device:
        dev = torch.device(self.device)
        if dev.type != "cuda":
            raise RuntimeError(
                f"device={self.device!r}: this build of awq_quantizer runs only on CUDA (B200, sm_100a); "
                "it has no CPU implementation")
        if not torch.cuda.is_available():
            raise RuntimeError("CUDA is not available and awq_quantizer (B200 build) has no CPU fallback")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        return dev

    def _layout(self, tensor: torch.Tensor):
        """(rows C, row length K, kernel group size, G, scale shape) per awq.py:297-323."""
        n = tensor.numel()
        if n < self.group_size:                                   # bypass, awq.py:297-300
            if tensor.dim() <= 1 or not self.per_channel:         # awq.py:147-149
                return 1, n, n, 1, ()
            C = tensor.shape[0]
            return C, n // C, n // C, 1, (C,)                     # awq.py:152-171
        if tensor.dim() <= 1:
            C, K = 1, n
        else:
            C = tensor.shape[0]
            K = n // C
        G = math.ceil(K / self.group_size)
        return C, K, self.group_size, G, (C, G)

    def _check_zero_point_mode(self):
        if self.zero_point == "percentile":
            # awq.py:189-190 passes 3 arguments to the 2-argument get_percentile_value -> TypeError
            raise TypeError("get_percentile_value() takes 2 positional arguments but 3 were given")

    # ------------------------------------------------------------------ awq.py:376-433
    def quantize(self, tensor: torch.Tensor, *, activations: Optional[torch.Tensor] = None,
                 pack: bool = False, keep_unpacked: bool = True) -> Dict[str, torch.Tensor]:
        if not isinstance(tensor, torch.Tensor):
            raise ValueError(f"Expected torch.Tensor, got {type(tensor)}")
        if not tensor.is_floating_point():
            raise ValueError(f"Expected floating point tensor, got {tensor.dtype}")
        dev = self._cuda_device()
        self._check_zero_point_mode()
        if tensor.numel() == 0:
            raise RuntimeError("cannot quantize an empty tensor (reference: min() of an empty tensor)")
        if activations is not None and self.scale_method != "mse":
            self.logger.warning("activations are ignored: the activation-aware search belongs to scale_method='mse'")
            activations = None
        if activations is not None:
            from .search import quantize_with_search
            return quantize_with_search(self, tensor, activations, dev, pack=pack)

        keep_unpacked = keep_unpacked or not pack
        routed = self._quantize_host_pipelined(tensor, dev, pack, keep_unpacked)
        if routed is not None:
            return routed
        w = tensor.to(dev, non_blocking=True).contiguous()
        out = self._quantize_device(w, pack=pack, unpacked=keep_unpacked)
        host = self._to_host({k: v for k, v in out.items() if v is not None}, dev)
        result = {
            "tensor_q": host["tensor_q"] if keep_unpacked else None,
            "scales": host["scales"],
            "zero_points": host["zero_points"],
            "bits": torch.tensor(self.bits, dtype=torch.int32),
            "group_size": torch.tensor(self.group_size, dtype=torch.int32),
            "symmetric": torch.tensor(self.symmetric, dtype=torch.bool),
        }
        if pack:
            result["qweight"] = host["qweight"]
            result["qzeros"] = host["qzeros"]
        if not keep_unpacked:
            del result["tensor_q"]
        return result

    _PIPELINE_MIN_ELEMS = 1 << 20

    def _quantize_host_pipelined(self, tensor: torch.Tensor, dev: torch.device, pack: bool,
                                 keep_unpacked: bool) -> Optional[Dict[str, torch.Tensor]]:
        """A large host tensor whose rows are whole groups does not make the upload -> kernel -> download round
        trip of awq.py:402-412 one after the other: it streams through the native gather pipeline (this thread's
        pipe: pinned rings, H2D / K1 / D2H overlapped).  None = not applicable, take the plain path."""
        from .arena import HostArena, arena_eligible, quantize_arena, short_row_len
        if tensor.device.type != "cpu" or tensor.numel() < self._PIPELINE_MIN_ELEMS:
            return None
        shape = tuple(tensor.shape)
        row_len = 0
        if not arena_eligible(shape, tensor.dtype, self.group_size, self.bits):
            row_len = short_row_len(shape, tensor.dtype, self.group_size, self.bits)
            if not row_len:
                return None
        src = {"t": tensor}
        r = quantize_arena(HostArena.for_tensors(src), bits=self.bits, group_size=self.group_size,
                           symmetric=self.symmetric, arith=self.arith, device=dev, packed=pack, unpacked=keep_unpacked,
                           want_zero_points=True, sources=src, pin_results=False, row_len=row_len)["t"]
        result = {"tensor_q": r.get("tensor_q"), "scales": r["scales"], "zero_points": r["zero_points"],
                  "bits": r["bits"], "group_size": r["group_size"], "symmetric": r["symmetric"]}
        if pack:
            result["qweight"], result["qzeros"] = r["qweight"], r["qzeros"]
        if not keep_unpacked:
            del result["tensor_q"]
        return result

    @staticmethod
    def _to_host(tensors: Dict[str, torch.Tensor], dev: torch.device) -> Dict[str, torch.Tensor]:
        """device -> host; the results are ordinary CPU tensors (awq.py:410-412 returns CPU tensors).
        Small results go through pinned buffers (torch's pinned allocator re-uses small blocks) with one stream
        sync for all of them; large ones are copied into pageable memory directly: a fresh pinned allocation
        costs 2.3 GB/s on this box, the driver's staged pageable copy runs at ~17 GB/s, and results that the
        caller keeps are never returned to the pinned cache."""
        out, big = {}, {}
        for k, v in tensors.items():
            if v.numel() * v.element_size() > (1 << 20):
                big[k] = v
                continue
            h = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
            h.copy_(v, non_blocking=True)
            out[k] = h
        for k, v in big.items():
            out[k] = v.cpu()
        torch.cuda.current_stream(dev).synchronize()
        return out

    def _quantize_device(self, w: torch.Tensor, *, pack: bool = False, unpacked: bool = True,
                         col_scale: Optional[torch.Tensor] = None,
                         arith: Optional[str] = None) -> Dict[str, torch.Tensor]:
        """Device-resident core: ``w`` is a contiguous CUDA tensor; returns CUDA tensors."""
        dev = w.device
        C, K, g, G, scale_shape = self._layout(w)
        arith = self.arith if arith is None else arith
        per = 32 // self.bits
        q = torch.empty(w.shape, dtype=torch.int32, device=dev) if unpacked else None
        scales = torch.empty((C, G), dtype=torch.float16, device=dev)
        zps = torch.empty((C, G), dtype=torch.int32, device=dev)
        qweight = torch.empty((C, -(-K // per)), dtype=torch.int32, device=dev) if pack else None
        qzeros = torch.empty((C, -(-G // per)), dtype=torch.int32, device=dev) if pack else None
        N.check(N.lib().awqk_group_quant(
            N.ptr(w), N.dtype_code(w.dtype), C, K, g, self.bits, int(self.symmetric),
            N.ARITH_FP32 if arith == "fp32" else N.ARITH_NATIVE,
            N.ptr(q), N.ptr(qweight), N.ptr(scales), N.ptr(zps), N.ptr(qzeros), N.ptr(col_scale),
            N.stream_ptr(dev)), "awqk_group_quant")
        out = {"tensor_q": q, "scales": scales.reshape(scale_shape), "zero_points": zps.reshape(scale_shape)}
        if pack:
            out["qweight"] = qweight
            out["qzeros"] = qzeros
        return out

    def _pin_now(self) -> bool:
        pin = self._models_done > 0 if self.pin_results is None else bool(self.pin_results)
        self._models_done += 1
        return pin

    # ------------------------------------------------------------------ awq.py:435-457
    def quantize_model(self, tensors, *, pack: bool = False, chunk_bytes: int = 32 << 20, pipeline: bool = True,
                       activations: Optional[Dict[str, torch.Tensor]] = None, keep_unpacked: Optional[bool] = None,
                       _pin: Optional[bool] = None):
        """dict-in / dict-out; a tensor that raises is logged and skipped (awq.py:453-455).

        ``pack=False`` (default): the reference's result layout per tensor (``tensor_q`` int32, ``scales``
        fp16, ``zero_points`` int32 ...).  ``pack=True``: packed results (``qweight`` / ``qzeros`` /
        ``scales``).  Either way every CPU tensor whose rows are whole groups (and, for the flat arena,
        whole packed-zero words) goes through ONE flat arena per dtype and the chunked H2D -> K1 -> D2H
        pipeline (quantization/arena.py) instead of a per-tensor upload / kernel / download sequence;
        the results are then views of pinned host arenas.  ``pipeline=False`` forces the per-tensor loop.
        ``tensors`` may already be a ``HostArena`` (zero-copy).

        ``activations`` (name -> calibration activations [tokens, in_features]) turns on the activation-aware
        alpha search for those tensors (quantization/search.py: streamed upload / search / download); all
        other tensors take the paths above.  With ``pack`` the results carry ``tensor_q`` / ``zero_points`` only
        if ``keep_unpacked`` is true (4 more bytes per element over PCIe)."""
        pin = self._pin_now() if _pin is None else _pin          # one decision per model
        if activations and self.scale_method != "mse":
            self.logger.warning("activations are ignored: the activation-aware search belongs to scale_method='mse'")
            activations = None
        if activations:
            from .search import quantize_model_with_search
            dev = self._cuda_device()
            self._check_zero_point_mode()
            want = {n: t for n, t in tensors.items() if n in activations}
            rest = {n: t for n, t in tensors.items() if n not in want}
            # the tensors without activations (embeddings, norms ...) go through the gather pipeline on a helper
            # thread WHILE the searched ones stream through the tensor-core path: their K1 launches are HBM bound
            # and slip in between the search kernels, their host copies use otherwise idle cores
            other: Dict[str, dict] = {}
            side_err: list = []

            def run_rest():
                try:
                    torch.cuda.set_device(dev)
                    other.update(self.quantize_model(rest, pack=pack, chunk_bytes=chunk_bytes, pipeline=pipeline,
                                                     keep_unpacked=keep_unpacked, _pin=pin))
                except BaseException as e:                       # reported below, on the caller's thread
                    side_err.append(e)

            side = None
            if rest:
                # a persistent worker: the gather pipeline's pinned rings belong to the thread that uses them
                if getattr(self, "_side_pool", None) is None:
                    from concurrent.futures import ThreadPoolExecutor
                    self._side_pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix="awq-rest")
                side = self._side_pool.submit(run_rest)
            searched = {}
            try:
                searched = quantize_model_with_search(self, want, activations, dev, pack=pack, keep_unpacked=keep_unpacked,
                                                      pin_results=bool(self.pin_results))
            except Exception as e:
                # the streamed model-level search failed as a whole (e.g. one bad shape, out of memory): search tensor
                # by tensor instead, so that one offender cannot silently turn AWQ scaling off for all the others
                self.logger.error(f"Streamed activation-aware search failed ({e}); searching tensor by tensor")
                if side is not None:
                    side.result()
                    side = None
                torch.cuda.synchronize(dev)
                keep = (not pack) if keep_unpacked is None else bool(keep_unpacked or not pack)
                for name, t in want.items():
                    try:
                        r = self.quantize(t, activations=activations[name], pack=pack)
                        if not keep:
                            r.pop("tensor_q", None)
                        searched[name] = r
                    except Exception as e2:
                        self.logger.error(f"Activation-aware search failed for {name}: {e2}; quantizing it without "
                                          f"AWQ scaling (no 'awq_scale' in its result)")
            if side is not None:
                side.result()
            if side_err:
                self.logger.error(f"Quantization of the tensors without activations failed: {side_err[0]}")
            missed = {n: t for n, t in want.items() if n not in searched}
            if missed:
                other.update(self.quantize_model(missed, pack=pack, chunk_bytes=chunk_bytes, pipeline=pipeline,
                                                 keep_unpacked=keep_unpacked, _pin=pin))
            return {n: (searched[n] if n in searched else other[n]) for n in tensors if n in searched or n in other}
        if not pack:
            from .arena import HostArena, pipe_eligible, quantize_arena
            quantized, rest = {}, tensors
            if pipeline and self.zero_point != "percentile" and torch.cuda.is_available() and \
                    torch.device(self.device).type == "cuda":
                dev = self._cuda_device()
                flat = {n: t for n, t in tensors.items()
                        if isinstance(t, torch.Tensor) and t.device.type == "cpu" and t.numel() >= self.group_size
                        and pipe_eligible(tuple(t.shape), t.dtype, self.group_size, self.bits)}
                if flat:
                    try:
                        res = quantize_arena(HostArena.for_tensors(flat), bits=self.bits, group_size=self.group_size,
                                             symmetric=self.symmetric, arith=self.arith, device=dev,
                                             chunk_bytes=chunk_bytes, packed=False, unpacked=True, sources=flat,
                                             pin_results=pin)
                        for name in flat:
                            self.logger.info(f"Successfully quantized tensor: {name}")
                        quantized.update(res)
                        rest = {n: t for n, t in tensors.items() if n not in flat}
                    except Exception as e:
                        self.logger.error(f"Pipelined quantization failed ({e}); using the per-tensor path")
            for name, tensor in rest.items():
                try:
                    self.logger.info(f"Quantizing tensor: {name}")
                    quantized[name] = self.quantize(tensor)
                    self.logger.info(f"Successfully quantized tensor: {name}")
                except Exception as e:
                    self.logger.error(f"Error quantizing tensor: {name}, error: {e}")
                    continue
            return {n: quantized[n] for n in tensors if n in quantized}       # input order, like the reference

        from .arena import (HostArena, arena_eligible, pipe_eligible, quantize_arena, quantize_rows_pipelined,
                            short_row_len, sync_pipe)
        dev = self._cuda_device()
        self._check_zero_point_mode()
        quantized = {}
        keep = bool(keep_unpacked)                     # also return the reference's tensor_q / zero_points
        if isinstance(tensors, HostArena):
            arena, singles = tensors, {}
            bad = [n for n, (shp, dt) in arena.specs.items() if not arena_eligible(shp, dt, self.group_size, self.bits)]
            if bad:
                raise ValueError(f"HostArena holds tensors that need the per-tensor path: {bad[:3]}")
        else:
            flat, singles = {}, {}
            for name, t in tensors.items():
                if isinstance(t, torch.Tensor) and t.device.type == "cpu" and \
                        arena_eligible(tuple(t.shape), t.dtype, self.group_size, self.bits):
                    flat[name] = t
                else:
                    singles[name] = t
            arena = HostArena.for_tensors(flat) if flat else None
        rest = {}

        def guarded(names, fn):
            """one pipelined section; on failure the pipe is drained BEFORE its half-written buffers are dropped
            (asynchronous D2H copies may still target them) and the tensors take the per-tensor path"""
            try:
                quantized.update(fn())
            except Exception as e:
                self.logger.error(f"Pipelined quantization failed ({e}); using the per-tensor path for {len(names)} tensors")
                try:
                    sync_pipe(dev, chunk_bytes)
                except Exception:
                    torch.cuda.synchronize(dev)
                for n in names:
                    quantized.pop(n, None)
                    rest[n] = names[n]

        if arena is not None:
            src = None if isinstance(tensors, HostArena) else flat
            guarded(src if src is not None else {n: arena.views[n] for n in arena.specs},
                    lambda: quantize_arena(arena, bits=self.bits, group_size=self.group_size,
                                           symmetric=self.symmetric, arith=self.arith, device=dev,
                                           chunk_bytes=chunk_bytes, sync=False, unpacked=keep, want_zero_points=keep,
                                           sources=src, pin_results=isinstance(tensors, HostArena) or pin))
        # rows of 1 / 2 / 4 groups (K = 512 at g = 128): gather pipeline too, one class per row length
        short: Dict[int, Dict[str, torch.Tensor]] = {}
        for name, tensor in list(singles.items()):
            # (a tensor that is already pinned takes the zero-copy, asynchronous row pipeline below)
            if isinstance(tensor, torch.Tensor) and tensor.device.type == "cpu" and not tensor.is_pinned():
                k = short_row_len(tuple(tensor.shape), tensor.dtype, self.group_size, self.bits)
                if k:
                    short.setdefault(k, {})[name] = singles.pop(name)
        for k, group in short.items():
            guarded(group, lambda k=k, group=group: quantize_arena(
                HostArena.for_tensors(group), bits=self.bits, group_size=self.group_size, symmetric=self.symmetric,
                arith=self.arith, device=dev, chunk_bytes=chunk_bytes, sync=False, unpacked=keep,
                want_zero_points=keep, sources=group, pin_results=pin, row_len=k))
        for name, tensor in singles.items():           # rows of whole groups: same pipeline, chunked by rows
            if not keep and isinstance(tensor, torch.Tensor) and tensor.device.type == "cpu" and \
                    pipe_eligible(tuple(tensor.shape), tensor.dtype, self.group_size, self.bits):
                guarded({name: tensor}, lambda name=name, tensor=tensor: {name: quantize_rows_pipelined(
                    tensor, bits=self.bits, group_size=self.group_size, symmetric=self.symmetric, arith=self.arith,
                    device=dev, chunk_bytes=chunk_bytes, sync=False)})
            else:
                rest[name] = tensor
        try:
            sync_pipe(dev, chunk_bytes)
        except Exception as e:                         # an asynchronous failure surfaces here: nothing queued is trusted
            self.logger.error(f"Pipelined quantization failed at the final synchronisation ({e}); using the per-tensor path")
            torch.cuda.synchronize(dev)
            for n in list(quantized):
                quantized.pop(n)
            rest = dict(tensors.items()) if not isinstance(tensors, HostArena) else {n: arena.views[n] for n in arena.specs}
        for r in quantized.values():
            r.pop("_keepalive", None)
        for name, tensor in rest.items():              # ragged rows, odd group sizes, fp64, numel < group_size ...
            try:
                quantized[name] = self.quantize(tensor, pack=True, keep_unpacked=keep)
            except Exception as e:
                self.logger.error(f"Error quantizing tensor: {name}, error: {e}")
        return quantized

    # ------------------------------------------------------------------ awq.py:459-539
    def dequantize(self, quantized_tensor: Dict[str, torch.Tensor]) -> torch.Tensor:
        dev = self._cuda_device()
        tensor_q = quantized_tensor["tensor_q"]
        scales = quantized_tensor["scales"]
        zero_points = quantized_tensor["zero_points"]
        group_size = int(quantized_tensor["group_size"].item())
        if scales.dim() != 2:
            # the reference indexes scales[c, g] (awq.py:516) -> IndexError on bypass layouts
            raise IndexError(f"too many indices for tensor of dimension {scales.dim()}")
        if tensor_q.numel() == 0:
            return torch.zeros(tensor_q.shape, dtype=torch.float32)
        if tensor_q.dim() <= 1:
            C, K = 1, tensor_q.numel()
        else:
            C, K = tensor_q.shape[0], tensor_q.numel() // tensor_q.shape[0]
        G = math.ceil(K / group_size)
        if tuple(scales.shape) != (C, G) or tuple(zero_points.shape) != (C, G):
            raise IndexError(f"scales/zero_points shape {tuple(scales.shape)} does not match {(C, G)}")
        q = tensor_q.to(dev, dtype=torch.int32, non_blocking=True).contiguous()
        s = scales.to(dev, dtype=torch.float16, non_blocking=True).contiguous()
        z = zero_points.to(dev, dtype=torch.int32, non_blocking=True).contiguous()
        out = torch.empty(tensor_q.shape, dtype=torch.float32, device=dev)
        N.check(N.lib().awqk_dequant(N.ptr(q), N.ptr(s), N.ptr(z), C, K, group_size, N.ptr(out),
                                     N.stream_ptr(dev)), "awqk_dequant")
        return out.cpu()

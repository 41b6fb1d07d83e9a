"""Deterministic synthetic inputs shared by the golden generator, the parity
tests, smoke() and bench.py.  numpy PCG64 streams (stable for a given numpy
version; every golden file also stores a digest of its inputs so drift is
detected instead of silently mis-compared)."""
from __future__ import annotations

import hashlib
import zlib

import numpy as np
import torch

DTYPES = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32, "fp64": torch.float64}


def seed_of(*parts) -> int:
    return zlib.crc32("/".join(str(p) for p in parts).encode()) ^ 0xA11CE


def weights(shape, dtype, seed: int, std: float = 0.02, offset: float = 0.0) -> torch.Tensor:
    """N(offset, std^2) in fp32, cast to dtype."""
    rng = np.random.Generator(np.random.PCG64(seed))
    a = rng.standard_normal(size=tuple(shape), dtype=np.float32) * np.float32(std) + np.float32(offset)
    t = torch.from_numpy(np.ascontiguousarray(a)).reshape(tuple(shape))
    return t.to(DTYPES[dtype] if isinstance(dtype, str) else dtype)


def activations(T: int, K: int, dtype, seed: int) -> torch.Tensor:
    """X[t,k] = N(0,1) * c_k, c_k = exp(N(0,1)) -- per-channel gain so that the
    alpha grid matters."""
    rng = np.random.Generator(np.random.PCG64(seed))
    gain = np.exp(rng.standard_normal(size=(K,), dtype=np.float32))
    a = rng.standard_normal(size=(T, K), dtype=np.float32) * gain[None, :]
    return torch.from_numpy(a).to(DTYPES[dtype] if isinstance(dtype, str) else dtype)


def raw_bytes(t: torch.Tensor) -> bytes:
    t = t.detach().cpu().contiguous()
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).numpy().tobytes()
    if t.dtype == torch.bool:
        return t.to(torch.uint8).numpy().tobytes()
    return t.numpy().tobytes()


def digest(*tensors) -> str:
    h = hashlib.sha256()
    for t in tensors:
        h.update(str(tuple(t.shape)).encode())
        h.update(str(t.dtype).encode())
        h.update(raw_bytes(t))
    return h.hexdigest()


def to_np(t: torch.Tensor) -> np.ndarray:
    """Storage form for npz files (bf16 as int16 bit pattern)."""
    t = t.detach().cpu().contiguous()
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).numpy()
    return t.numpy()


def from_np(a: np.ndarray, dtype) -> torch.Tensor:
    dtype = DTYPES[dtype] if isinstance(dtype, str) else dtype
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype == torch.bfloat16:
        return t.view(torch.bfloat16)
    return t

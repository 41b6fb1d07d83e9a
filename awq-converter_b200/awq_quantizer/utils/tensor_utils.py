"""Tensor helpers on the hot path.  Mirrors the reference's ``utils/tensor_utils.py``:

* ``convert_bf16_to_fp16(tensor)`` (tensor_utils.py:10-22): bf16 -> fp16 (RNE) on the B200 kernel
  ``awqk_bf16_to_fp16``; any other dtype is returned unchanged (the same object).  The result lives
  on the device of the input (a CPU tensor is staged through the GPU: there is no CPU arithmetic).
* the safetensors file-selection helpers (tensor_utils.py:203-313) -- pure path logic used by the
  loader.
"""
from __future__ import annotations

import os
from typing import List

import torch

from .. import _native as N


def convert_bf16_to_fp16(tensor: torch.Tensor, device=None) -> torch.Tensor:
    if tensor.dtype != torch.bfloat16:
        return tensor
    if not torch.cuda.is_available():
        raise RuntimeError("CUDA is not available and awq_quantizer (B200 build) has no CPU fallback")
    src_device = tensor.device
    if src_device.type == "cuda":
        dev = src_device
    else:
        dev = torch.device(device if device is not None else "cuda")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
    x = tensor.to(dev, non_blocking=True).contiguous()
    out = torch.empty(x.shape, dtype=torch.float16, device=dev)
    N.check(N.lib().awqk_bf16_to_fp16(N.ptr(x), N.ptr(out), x.numel(), N.stream_ptr(dev)),
            "awqk_bf16_to_fp16")
    return out if src_device.type == "cuda" else out.to(src_device)


def is_consolidated_file(file_path: str) -> bool:
    return file_path.endswith(".safetensors") and "consolidated" in os.path.basename(file_path).lower()


def filter_safetensor_files(file_paths: List[str]) -> List[str]:
    """shards win over consolidated files (tensor_utils.py:203-238)"""
    st = [p for p in file_paths if p.endswith(".safetensors")]
    shards = [p for p in st if not is_consolidated_file(p)]
    return shards if shards else st


def filter_consolidated_files(files: List[str]) -> List[str]:
    """tensor_utils.py:281-313: a single file is kept as is; otherwise sorted shards, else sorted
    consolidated files"""
    if len(files) <= 1:
        return files
    shards = [f for f in files if not is_consolidated_file(f)]
    return sorted(shards) if shards else sorted(f for f in files if is_consolidated_file(f))


def get_model_files(model_path: str) -> List[str]:
    """tensor_utils.py:258-278"""
    if os.path.isfile(model_path):
        return [model_path] if model_path.endswith(".safetensors") else []
    found = []
    for root, _, names in os.walk(model_path):
        found.extend(os.path.join(root, n) for n in names if n.endswith(".safetensors"))
    return filter_safetensor_files(found)

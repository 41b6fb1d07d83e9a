"""Interop export: a packed quantize() result -> the AutoAWQ / vLLM "GEMM" checkpoint layout
(SURVEY.md section 8f, rank 4).  The reference's closest equivalent is its non-functional
examples/load_quantized_model.py.  Re-layout only; runs in csrc/awqk_export.cu."""
from __future__ import annotations

from typing import Dict

import torch

from .. import _native as N


def to_autoawq_gemm(result: Dict[str, torch.Tensor], device="cuda") -> Dict[str, torch.Tensor]:
    """``result`` = ``AWQQuantizer.quantize(w, pack=True)`` (or ``quantize_model(..., pack=True)[name]``
    plus ``zero_points``) of a 2-D int4 weight [C, K].  Returns CPU tensors ``qweight`` int32 [K, C/8],
    ``qzeros`` int32 [G, C/8], ``scales`` fp16 [G, C] -- ``w ~ (q - z) * s`` -- plus ``awq_scale`` if searched."""
    if int(result["bits"]) != 4:
        raise ValueError("the AutoAWQ GEMM layout is int4 only")
    if "zero_points" not in result or "qweight" not in result:
        raise ValueError("need a packed result with zero_points (quantize(..., pack=True))")
    dev = torch.device(device)
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    qw = result["qweight"].to(dev, non_blocking=True).contiguous()
    zp = result["zero_points"].to(dev, non_blocking=True).contiguous()
    sc = result["scales"].to(dev, non_blocking=True).contiguous()
    if zp.dim() != 2:
        raise ValueError("need [C, G] zero points (2-D weight)")
    C, G = zp.shape
    K = qw.shape[1] * 8
    out_q = torch.empty((K, C // 8), dtype=torch.int32, device=dev)
    out_z = torch.empty((G, C // 8), dtype=torch.int32, device=dev)
    out_s = torch.empty((G, C), dtype=torch.float16, device=dev)
    N.check(N.lib().awqk_export_autoawq(N.ptr(qw), N.ptr(zp), N.ptr(sc), C, K, G, int(bool(result["symmetric"])),
                                        N.ptr(out_q), N.ptr(out_z), N.ptr(out_s), N.stream_ptr(dev)),
            "awqk_export_autoawq")
    out = {"qweight": out_q.cpu(), "qzeros": out_z.cpu(), "scales": out_s.cpu()}
    if "awq_scale" in result:
        out["awq_scale"] = result["awq_scale"]
    return out

"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/awqk.h
declares; argument validation that needs no GPU."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "awqk.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"AWQK_API\s+[\w\s\*]+?\b(awqk_\w+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    names = declared_functions()
    for must in ("awqk_group_quant", "awqk_dequant", "awqk_bf16_to_fp16", "awqk_sqerr_gemm",
                 "awqk_fakequant_delta", "awqk_abs_colsum", "awqk_alpha_grid", "awqk_pipe_quant_host"):
        assert must in names


def test_library_exports_every_declared_symbol(native_lib):
    from awq_quantizer import _native
    names = declared_functions()
    assert sorted(_native.SIGNATURES) == names, "ctypes table out of sync with include/awqk.h"
    for n in names:
        assert hasattr(native_lib, n), f"libawqk.so does not export {n}"
    out = subprocess.run(["nm", "-D", "--defined-only", _native.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (awqk_\w+)", out))
    assert exported == set(names), exported ^ set(names)


def test_library_is_sm100a_only(native_lib):
    from awq_quantizer import _native
    out = subprocess.run(["cuobjdump", "-lelf", _native.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_version_and_error_strings(native_lib):
    assert native_lib.awqk_version() == 210
    assert native_lib.awqk_error_string(0) == b"ok"
    assert b"argument" in native_lib.awqk_error_string(-1)
    assert native_lib.awqk_error_string(-99) == b"unknown error"


def test_argument_validation_without_gpu(native_lib):
    L = native_lib
    # null weight pointer / bad bits / bad sizes are rejected before any CUDA call
    assert L.awqk_group_quant(None, 0, 4, 128, 128, 4, 0, 0, None, None, None, None, None, None, None) == -1
    assert L.awqk_group_quant_path(0, 4, 128, 128, 3, 0, None) == -1
    assert L.awqk_group_quant_path(0, 0, 128, 128, 4, 0, None) == -1
    assert L.awqk_group_quant_path(9, 4, 128, 128, 4, 0, None) == -1
    # path selection is pure host logic: flat path needs g in {32,64,128}, K % g == 0, 16-B base
    buf = ctypes.create_string_buffer(64)
    base = ctypes.addressof(buf)
    base += (-base) % 16
    assert L.awqk_group_quant_path(0, 4, 256, 128, 4, 0, base) == 1
    assert L.awqk_group_quant_path(0, 4, 300, 128, 4, 0, base) == 0
    assert L.awqk_group_quant_path(0, 4, 256, 100, 4, 0, base) == 0
    assert L.awqk_group_quant_path(3, 4, 256, 128, 4, 0, base) == 0
    assert L.awqk_group_quant_path(0, 4, 256, 128, 4, 0, base + 2) == 0
    assert L.awqk_bf16_to_fp16(None, None, 0, None) == 0
    assert L.awqk_bf16_to_fp16(None, None, -1, None) == -1
    assert L.awqk_dequant(None, None, None, 1, 1, 1, None, None) == -1


def test_host_copy_without_gpu(native_lib):
    """awqk_host_copy is plain host code (parallel memcpy): usable, and exact, without a device"""
    import torch
    from awq_quantizer import _native
    a = torch.arange(3 * (1 << 20) + 17, dtype=torch.int32)
    b = torch.zeros_like(a)
    _native.host_copy(b, a)
    assert torch.equal(a, b)
    c = torch.arange(600, dtype=torch.float32).reshape(200, 3)[:, :2]     # non-contiguous -> torch path
    d = torch.empty(200, 2)
    _native.host_copy(d, c)
    assert torch.equal(c, d)
    assert native_lib.awqk_host_copy(None, None, 16, 2) == -1

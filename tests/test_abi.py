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


def _plan(native_lib, shapes, g=128, unpacked=0, sms=148):
    n = len(shapes)
    C = (ctypes.c_int64 * n)(*[s[0] for s in shapes])
    K = (ctypes.c_int64 * n)(*[s[1] for s in shapes])
    summ = (ctypes.c_int64 * 5)()
    items = (ctypes.c_int64 * (4 * 32))()
    rc = native_lib.awqk_group_quant_batch_plan(C, K, n, g, unpacked, sms, summ, items, 32)
    return rc, list(summ), [tuple(items[4 * i:4 * i + 4]) for i in range(int(summ[3]))] if rc == 0 else []


def test_batch_plan_partitions_every_tensor(native_lib):
    """the unit plan of the column-slab launch (awqk_group_quant_batch): the launch items partition the rows of
    every tensor in order, short units only at the end of the tensor list, unit count consistent, and the modelled
    cost never above the best single-height plan"""
    import random
    rnd = random.Random(7)
    layer = [(4096, 4096), (1024, 4096), (1024, 4096), (4096, 4096), (14336, 4096), (14336, 4096), (4096, 14336)]
    cases = [[(8192, 28672)], [(28672, 8192)], [(1, 1024)], [(7, 2048), (9, 1024)], layer, layer * 4, [(4096, 4096)] * 31]
    for _ in range(40):
        cases.append([(rnd.randint(1, 20000), 1024 * rnd.randint(1, 28)) for _ in range(rnd.randint(1, 31))])
    for shapes in cases:
        for unpacked in (0, 1):
            rc, (r_main, r_tail, units, n_items, cost), items = _plan(native_lib, shapes, unpacked=unpacked)
            assert rc == 0, (rc, shapes)
            grid = 148 * (2 if unpacked else 3)
            assert r_main in (8, 16, 32, 64, 128) and r_tail in (0, 8, 16, 32, 64) and r_tail < r_main
            assert 1 <= n_items <= 32 and len(items) == n_items
            # partition: per tensor, consecutive row ranges starting at 0 and ending at C, in tensor order
            pos = {}
            order = []
            seen_short = False
            u = 0
            for t, r0, rows, r in items:
                assert r in (r_main, r_tail) and r > 0 and rows > 0
                assert r0 == pos.get(t, 0), (shapes, items)
                pos[t] = r0 + rows
                order.append(t)
                if r != r_main:
                    seen_short = True
                else:
                    assert not seen_short, "tall units after the short tail"
                u += (shapes[t][1] // 1024) * -(-rows // r)
            assert order == sorted(order)
            assert all(pos.get(t, 0) == c for t, (c, _) in enumerate(shapes)), (shapes, items)
            assert u == units
            # never worse than the best single-height plan under the same cost model
            single = min(-(-sum((k // 1024) * -(-c // r) for c, k in shapes) // grid) * (r + 2) for r in (8, 16, 32, 64, 128))
            assert cost <= single, (cost, single, shapes)
            # and close to the ideal (total rows x slabs / grid) for launches of at least a few rounds
            ideal = sum(c * (k // 1024) for c, k in shapes) / grid
            if ideal >= 256:
                assert cost <= 1.06 * ideal + 12, (cost, ideal, shapes)


def test_batch_plan_rejects_bad_arguments(native_lib):
    assert _plan(native_lib, [(16, 1536)])[0] == -4            # K % 1024 != 0: not a column-slab tensor
    assert _plan(native_lib, [(16, 1024)] * 32)[0] == -1       # more than 31 tensors per launch
    assert _plan(native_lib, [(16, 1024)], sms=0)[0] == -1
    assert _plan(native_lib, [(16, 1024)], g=48)[0] == -1


def test_quant_item_struct_layout_matches_header(tmp_path):
    """the ctypes mirror of awqk_quant_item (used by awqk_group_quant_batch) has the size and field offsets the C
    compiler gives the struct declared in include/awqk.h"""
    from awq_quantizer import _native
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "awqk.h"\n'
                   'int main(void) { printf("%zu", sizeof(awqk_quant_item));\n'
                   + "".join(f'  printf(" %zu", offsetof(awqk_quant_item, {f}));\n' for f, _ in _native.QuantItem._fields_)
                   + "  return 0; }\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [ctypes.sizeof(_native.QuantItem)] + [getattr(_native.QuantItem, f).offset for f, _ in _native.QuantItem._fields_]
    assert got == want, (got, want)

"""CPU: the oracle (oracle/awq_oracle.py) against the fixtures frozen from the reference itself
(tests/golden/make_golden.py).  This is what pins the oracle; the GPU parity tests then compare the
CUDA path with the oracle / the same fixtures."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import awq_oracle as O
from tests import datagen
from tests.golden import cases
from tests.util import assert_quant_equal, assert_same

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def small():
    return np.load(os.path.join(GOLD, "small.npz"))


@pytest.fixture(scope="module")
def special():
    return np.load(os.path.join(GOLD, "special.npz"))


def golden_result(npz, key):
    return {
        "tensor_q": torch.from_numpy(npz[key + "/tensor_q"]),
        "scales": torch.from_numpy(npz[key + "/scales"]).view(torch.float16),
        "zero_points": torch.from_numpy(npz[key + "/zero_points"]),
    }


def test_small_cases_match_reference(small):
    n = 0
    for c in cases.small_cases():
        key = cases.case_key(c)
        w = cases.case_input(c)
        assert bytes(small[key + "/in_digest"]).hex() == datagen.digest(w), f"input drift for {key}"
        got = O.group_quant_vec(w, c["bits"], c["group_size"], c["symmetric"], c["per_channel"])
        want = golden_result(small, key)
        assert_quant_equal(got, want, key)
        if key + "/dequant" in small.files:
            assert_same(O.dequant_vec({**want, "group_size": torch.tensor(c["group_size"])}),
                        torch.from_numpy(small[key + "/dequant"]), key + "/dequant")
        else:
            with pytest.raises(IndexError):
                O.dequant_vec({**want, "group_size": torch.tensor(c["group_size"])})
        n += 1
    assert n >= 400


def test_special_values_match_reference(special):
    for name in cases.special_inputs():
        for dt in ("bf16", "fp16", "fp32"):
            for sym in (False, True):
                key = f"{name}_{dt}_{'sym' if sym else 'asym'}"
                raw = special[key + "/input"]
                w = datagen.from_np(raw, dt).view(datagen.DTYPES[dt]) if dt == "fp16" else datagen.from_np(raw, dt)
                got = O.group_quant_vec(w, 4, 128, sym, True)
                want = golden_result(special, key)
                assert_quant_equal(got, want, key)
                assert_same(O.dequant_vec({**want, "group_size": torch.tensor(128)}),
                            torch.from_numpy(special[key + "/dequant"]), key + "/dequant")


def test_loop_port_equals_vectorised():
    """the group-at-a-time port (what --impl reference times) == the vectorised oracle"""
    for shape, dt, sym, g in [((6, 300), "bf16", False, 128), ((3, 256), "fp16", True, 64),
                              ((2, 5, 70), "fp32", False, 32), ((200,), "bf16", True, 128)]:
        w = datagen.weights(shape, dt, datagen.seed_of("loop", shape, dt), offset=0.1)
        assert_quant_equal(O.group_quant_loop(w, 4, g, sym, True), O.group_quant_vec(w, 4, g, sym, True),
                           f"{shape}/{dt}")


def test_medium_digests():
    with open(os.path.join(GOLD, "medium.json")) as f:
        med = json.load(f)
    for c in cases.MEDIUM_CASES:
        w = cases.medium_input(c)
        if c["convert_fp16"]:
            w = O.bf16_to_fp16(w)
        want = med[c["name"]]
        assert datagen.digest(w) == want["input"], "input drift " + c["name"]
        r = O.group_quant_vec(w, 4, 128, c["symmetric"], True)
        for k in ("tensor_q", "scales", "zero_points"):
            assert datagen.digest(r[k]) == want[k], (c["name"], k)
        assert datagen.digest(O.dequant_vec(r)) == want["dequant"], c["name"]


def test_bf16_to_fp16_exhaustive():
    gold = np.load(os.path.join(GOLD, "convert.npz"))["fp16_bits"]
    allbits = torch.arange(65536, dtype=torch.int32).to(torch.int16).view(torch.bfloat16)
    got = O.bf16_to_fp16(allbits)
    assert got.dtype == torch.float16
    assert_same(got, torch.from_numpy(gold).view(torch.float16), "bf16->fp16")
    f16 = torch.ones(3, dtype=torch.float16)
    assert O.bf16_to_fp16(f16) is f16                       # tensor_utils.py:20-22 identity


def test_fp32_arith_is_reference_on_float():
    w = datagen.weights((4, 512), "bf16", 7)
    a = O.group_quant_vec(w, 4, 128, False, True, arith="fp32")
    b = O.group_quant_vec(w.float(), 4, 128, False, True)
    assert_quant_equal(a, b)


def test_errors():
    with pytest.raises(ValueError):
        O.group_quant_vec(torch.zeros(4, 4, dtype=torch.int32))
    with pytest.raises(ValueError):
        O.group_quant_vec([1.0, 2.0])
    with pytest.raises(RuntimeError):
        O.group_quant_vec(torch.zeros(0, 5))


@pytest.mark.parametrize("bits,sym", [(4, False), (4, True), (8, False), (8, True)])
def test_pack_roundtrip(bits, sym):
    qmin, qmax = O.qrange(bits, sym)
    g = torch.Generator().manual_seed(bits * 2 + sym)
    for n in (1, 7, 8, 9, 64, 100):
        codes = torch.randint(qmin, qmax + 1, (5, n), generator=g, dtype=torch.int32)
        words = O.pack_rows_u32(codes, qmin, bits)
        per = 32 // bits
        assert words.shape == (5, -(-n // per)) and words.dtype == torch.int32
        assert torch.equal(O.unpack_rows_u32(words, n, qmin, bits), codes)
    # explicit little-endian nibble order: code i sits at bits [4i, 4i+4)
    codes = torch.arange(8, dtype=torch.int32).reshape(1, 8) + qmin
    if bits == 4:
        assert int(O.pack_rows_u32(codes, qmin, 4)[0, 0]) == 0x76543210


def test_pack_result_consistent_with_quantize():
    w = datagen.weights((8, 1000), "bf16", 11, offset=0.2)
    r = O.pack_result(O.group_quant_vec(w, 4, 128, False, True))
    assert r["qweight"].shape == (8, 125) and r["qzeros"].shape == (8, 1)
    assert torch.equal(O.unpack_rows_u32(r["qweight"], 1000, 0), r["tensor_q"])
    assert torch.equal(O.unpack_rows_u32(r["qzeros"], 8, 0), r["zero_points"])


def test_search_oracle_prefers_activation_aware_scaling():
    """sanity of the (unpinned) alpha-search definition: with strongly non-uniform channel gains the
    chosen alpha is > 0 and its error is below the alpha = 0 (plain quantization) error."""
    W = datagen.weights((64, 256), "bf16", 3)
    X = datagen.activations(128, 256, "bf16", 4)
    r = O.search_scales(W, X, 4, 128, False, n_grid=10)
    assert r["best_idx"] > 0 and r["err"][r["best_idx"]] < r["err"][0]
    assert abs(float(r["s_grid"][0].max()) - 1.0) < 1e-6          # alpha = 0 -> s == 1
    # error of alpha=0 equals plain quantization error
    dW = O.fake_quant_delta(W, torch.ones(256), 4, 128, False).double()
    ref = float(((X.double() @ dW.T) ** 2).mean())
    assert abs(ref - r["err"][0]) <= 1e-12 * max(1.0, ref)


def test_c_oracle_matches_goldens_and_python_oracle(small, special):
    """the plain-C restatement (oracle/awq_oracle.c, also the fast CPU baseline) against the same
    reference fixtures"""
    from oracle import c_oracle as CO
    n = 0
    for c in cases.small_cases():
        if c["dtype"] == "fp64":
            continue
        w = cases.case_input(c)
        if w.numel() < c["group_size"]:
            continue                                        # bypass layouts are host logic, not in the C core
        key = cases.case_key(c)
        got = CO.group_quant(w, c["bits"], c["group_size"], c["symmetric"], threads=2)
        want = golden_result(small, key)
        assert_quant_equal(got, {k: v.reshape(got[k].shape) for k, v in want.items()}, key)
        n += 1
    assert n > 300
    for name, master in cases.special_inputs().items():
        for dt in ("bf16", "fp16", "fp32"):
            w = master.to(datagen.DTYPES[dt])
            for sym in (False, True):
                key = f"{name}_{dt}_{'sym' if sym else 'asym'}"
                got = CO.group_quant(w, 4, 128, sym, threads=2)
                assert_quant_equal(got, golden_result(special, key), key)
                assert_same(CO.dequant(got), torch.from_numpy(special[key + "/dequant"]), key + "/dequant")
                assert_quant_equal(CO.group_quant(w, 4, 128, sym, arith="fp32", threads=2),
                                   O.group_quant_vec(w, 4, 128, sym, True, arith="fp32"), key + "/fp32")
    codes = torch.randint(0, 16, (7, 100), dtype=torch.int32)
    assert torch.equal(CO.pack_rows(codes, 0, 4), O.pack_rows_u32(codes, 0, 4))


def test_autoawq_layout_roundtrip():
    w = datagen.weights((64, 256), "bf16", 5)
    r = O.group_quant_vec(w, 4, 128, False, True)
    a = O.to_autoawq_gemm(r)
    assert a["qweight"].shape == (256, 8) and a["qzeros"].shape == (2, 8) and a["scales"].shape == (2, 64)
    assert torch.equal(O.from_autoawq_gemm(a["qweight"], 64).t(), r["tensor_q"])
    assert torch.equal(O.from_autoawq_gemm(a["qzeros"], 64).t(), r["zero_points"])
    # first word of row 0 holds output channels 0,2,4,6,1,3,5,7 of input feature 0
    q = r["tensor_q"][:, 0]
    want = sum(int(q[o]) << (4 * i) for i, o in enumerate(O.AWQ_ORDER))
    assert (int(a["qweight"][0, 0]) & 0xFFFFFFFF) == want

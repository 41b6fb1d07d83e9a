#!/usr/bin/env python
"""Generate the golden fixtures by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Imports ``awq_quantizer`` from /root/reference/src (nothing is copied), feeds it
the deterministic inputs of tests/datagen.py, and freezes what it returns:

* small.npz    -- per case: outputs of AWQQuantizer.quantize (+ dequantize when
                  it works) on the small/edge shapes of cases.py, for all
                  dtypes / symmetric / bits / group sizes;
* special.npz  -- NaN / inf / all-zero / constant / tie inputs AND outputs;
* medium.json  -- sha256 digests of outputs for the config-0 shapes
                  (test_quantization.py:54-63) and CLI-style bf16 / fp32 runs;
* convert.npz  -- bf16 -> fp16 conversion table (tensor_utils.py:10-22).

While generating it also asserts that oracle/awq_oracle.py reproduces every one
of these outputs bit-for-bit (vectorised and group-at-a-time forms) -- this is
the step that PINS the oracle.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

from awq_quantizer.quantization.awq import AWQQuantizer as RefQuantizer  # noqa: E402  (the reference)
from awq_quantizer.utils.tensor_utils import convert_bf16_to_fp16 as ref_convert  # noqa: E402

assert "/root/reference/" in sys.modules["awq_quantizer"].__file__, "must import the reference"

from oracle import awq_oracle as O  # noqa: E402
from tests import datagen  # noqa: E402
from tests.golden import cases  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
KEYS = ("tensor_q", "scales", "zero_points")


def same_f32(a, b):
    """bitwise equality except that any NaN matches any NaN"""
    return a.shape == b.shape and torch.equal(torch.nan_to_num(a, nan=1234.5), torch.nan_to_num(b, nan=1234.5)) \
        and torch.equal(torch.isnan(a), torch.isnan(b))


def ref_quant(w, bits, g, sym, pc):
    q = RefQuantizer(bits=bits, group_size=g, symmetric=sym, per_channel=pc, device="cpu",
                     logger_level="ERROR")
    return q, q.quantize(w)


def canon(t):
    """fp16 -> int16 bit pattern with every NaN mapped to 0x7E00 (NaN sign and
    payload are not part of the parity contract); other dtypes unchanged."""
    if t.dtype != torch.float16:
        return t
    bits = t.view(torch.int16).clone()
    bits[torch.isnan(t)] = 0x7E00
    return bits


def check_oracle(w, r, bits, g, sym, pc, loop=True, tag=""):
    o = O.group_quant_vec(w, bits, g, sym, pc)
    for k in KEYS + ("bits", "group_size", "symmetric"):
        assert o[k].dtype == r[k].dtype and o[k].shape == r[k].shape, (tag, k, o[k].dtype, r[k].dtype, o[k].shape, r[k].shape)
        assert torch.equal(canon(o[k]), canon(r[k])), (tag, k)
    if loop:
        l = O.group_quant_loop(w, bits, g, sym, pc)
        for k in KEYS:
            assert torch.equal(canon(l[k]), canon(r[k])), (tag, "loop", k)


def main():
    t0 = time.time()
    small = {}
    n = 0
    for c in cases.small_cases():
        key = cases.case_key(c)
        w = cases.case_input(c)
        qz, r = ref_quant(w, c["bits"], c["group_size"], c["symmetric"], c["per_channel"])
        check_oracle(w, r, c["bits"], c["group_size"], c["symmetric"], c["per_channel"], tag=key)
        small[key + "/in_digest"] = np.frombuffer(bytes.fromhex(datagen.digest(w)), dtype=np.uint8)
        for k in KEYS:
            small[key + "/" + k] = datagen.to_np(r[k].view(torch.int16) if r[k].dtype == torch.float16 else r[k])
        if r["scales"].dim() == 2:
            d = qz.dequantize(r)
            assert same_f32(O.dequant_vec(r), d), key
            small[key + "/dequant"] = d.numpy()
        else:
            try:
                qz.dequantize(r)
                raise AssertionError("reference dequantize unexpectedly worked on " + key)
            except IndexError:
                pass
        n += 1
    np.savez_compressed(os.path.join(OUT, "small.npz"), **small)
    print(f"small: {n} cases, {time.time() - t0:.1f}s")

    special = {}
    for name, master in cases.special_inputs().items():
        for dt in ("bf16", "fp16", "fp32"):
            w = master.to(datagen.DTYPES[dt])
            for sym in (False, True):
                key = f"{name}_{dt}_{'sym' if sym else 'asym'}"
                qz, r = ref_quant(w, 4, 128, sym, True)
                check_oracle(w, r, 4, 128, sym, True, tag=key)
                special[key + "/input"] = datagen.to_np(w.view(torch.int16) if dt == "fp16" else w)
                for k in KEYS:
                    special[key + "/" + k] = datagen.to_np(r[k].view(torch.int16) if r[k].dtype == torch.float16 else r[k])
                d = qz.dequantize(r)
                od = O.dequant_vec(r)
                assert same_f32(od, d), key
                special[key + "/dequant"] = d.numpy()
    np.savez_compressed(os.path.join(OUT, "special.npz"), **special)
    print(f"special: {len(special)} arrays, {time.time() - t0:.1f}s")

    medium = {}
    for c in cases.MEDIUM_CASES:
        w = cases.medium_input(c)
        if c["convert_fp16"]:
            w2 = ref_convert(w)                       # test_quantization.py:132
            assert torch.equal(O.bf16_to_fp16(w).view(torch.int16), w2.view(torch.int16))
            w = w2
        t1 = time.time()
        qz, r = ref_quant(w, 4, 128, c["symmetric"], True)
        dt_ref = time.time() - t1
        check_oracle(w, r, 4, 128, c["symmetric"], True, loop=False, tag=c["name"])
        d = qz.dequantize(r)
        assert same_f32(O.dequant_vec(r), d)
        medium[c["name"]] = {
            "input": datagen.digest(w), "tensor_q": datagen.digest(r["tensor_q"]),
            "scales": datagen.digest(r["scales"]), "zero_points": datagen.digest(r["zero_points"]),
            "dequant": datagen.digest(d), "ref_seconds_here": round(dt_ref, 3),
            "groups": int(r["scales"].numel()),
        }
        print("medium", c["name"], f"{dt_ref:.2f}s")
    with open(os.path.join(OUT, "medium.json"), "w") as f:
        json.dump(medium, f, indent=1, sort_keys=True)

    # bf16 -> fp16: every finite-exponent class plus specials, 4096 patterns + all 65536 exhaustively digested
    allbits = torch.arange(65536, dtype=torch.int32).to(torch.int16).view(torch.bfloat16)
    conv = ref_convert(allbits)
    assert conv.dtype == torch.float16
    assert torch.equal(O.bf16_to_fp16(allbits).view(torch.int16), conv.view(torch.int16))
    same = ref_convert(conv)
    assert same is conv                               # non-bf16 returned unchanged (same object)
    np.savez_compressed(os.path.join(OUT, "convert.npz"), fp16_bits=conv.view(torch.int16).numpy())
    print(f"done in {time.time() - t0:.1f}s; torch {torch.__version__}")


if __name__ == "__main__":
    main()

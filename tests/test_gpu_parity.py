"""GPU (B200): the CUDA path, called through the reference-facing Python mirror and the C ABI,
against (a) the fixtures frozen from the reference, (b) the pinned oracle on seeded inputs,
(c) size-independent properties at BASELINE.json's full micro-bench size.

Bar: bit-exact for tensor_q / zero_points / qweight / qzeros and for the fp16 scales (NaN payloads
excepted)."""
import concurrent.futures as cf
import json
import os

import numpy as np
import pytest
import torch

from oracle import awq_oracle as O
from tests import datagen
from tests.golden import cases
from tests.util import assert_quant_equal, assert_same

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def mk(**kw):
    from awq_quantizer.quantization import AWQQuantizer
    kw.setdefault("logger_level", "ERROR")
    kw.setdefault("device", "cuda:0")
    return AWQQuantizer(**kw)


def golden_result(npz, key):
    return {
        "tensor_q": torch.from_numpy(npz[key + "/tensor_q"]),
        "scales": torch.from_numpy(npz[key + "/scales"]).view(torch.float16),
        "zero_points": torch.from_numpy(npz[key + "/zero_points"]),
    }


def test_small_cases_vs_reference_golden(native_lib, cuda_device):
    small = np.load(os.path.join(GOLD, "small.npz"))
    quantizers = {}
    n = 0
    for c in cases.small_cases():
        key = cases.case_key(c)
        cfg = (c["bits"], c["group_size"], c["symmetric"], c["per_channel"])
        if cfg not in quantizers:
            quantizers[cfg] = mk(bits=c["bits"], group_size=c["group_size"], symmetric=c["symmetric"],
                                 per_channel=c["per_channel"])
        qz = quantizers[cfg]
        w = cases.case_input(c)
        got = qz.quantize(w)
        want = golden_result(small, key)
        assert_quant_equal(got, want, key)
        assert got["tensor_q"].device.type == "cpu" and got["scales"].device.type == "cpu"
        assert int(got["bits"]) == c["bits"] and int(got["group_size"]) == c["group_size"]
        assert got["bits"].dtype == torch.int32 and got["symmetric"].dtype == torch.bool
        assert bool(got["symmetric"]) == c["symmetric"]
        if key + "/dequant" in small.files:
            assert_same(qz.dequantize(got), torch.from_numpy(small[key + "/dequant"]), key + "/dequant")
        else:
            with pytest.raises(IndexError):
                qz.dequantize(got)
        n += 1
    assert n >= 400


def test_special_values_vs_reference_golden(native_lib, cuda_device):
    special = np.load(os.path.join(GOLD, "special.npz"))
    for name in cases.special_inputs():
        for dt in ("bf16", "fp16", "fp32"):
            for sym in (False, True):
                key = f"{name}_{dt}_{'sym' if sym else 'asym'}"
                raw = special[key + "/input"]
                w = datagen.from_np(raw, dt).view(datagen.DTYPES[dt]) if dt == "fp16" else datagen.from_np(raw, dt)
                qz = mk(symmetric=sym)
                got = qz.quantize(w)
                assert_quant_equal(got, golden_result(special, key), key)
                assert_same(qz.dequantize(got), torch.from_numpy(special[key + "/dequant"]), key + "/dequant")


def test_config0_shapes_digests(native_lib, cuda_device):
    """test_quantization.py:54-63,132,136-145 shapes; digests of the reference's own outputs."""
    from awq_quantizer.utils.tensor_utils import convert_bf16_to_fp16
    with open(os.path.join(GOLD, "medium.json")) as f:
        med = json.load(f)
    for c in cases.MEDIUM_CASES:
        w = cases.medium_input(c)
        if c["convert_fp16"]:
            w = convert_bf16_to_fp16(w)
            assert w.dtype == torch.float16
        want = med[c["name"]]
        assert datagen.digest(w) == want["input"], "converted input differs " + c["name"]
        qz = mk(symmetric=c["symmetric"])
        r = qz.quantize(w)
        for k in ("tensor_q", "scales", "zero_points"):
            assert datagen.digest(r[k]) == want[k], (c["name"], k)
        assert datagen.digest(qz.dequantize(r)) == want["dequant"], c["name"]
    with pytest.raises(ValueError):                     # the 100x100 int32 tensor must be rejected
        mk().quantize(torch.zeros(100, 100, dtype=torch.int32))


@pytest.mark.parametrize("dt", ["bf16", "fp16", "fp32"])
@pytest.mark.parametrize("sym", [False, True])
@pytest.mark.parametrize("g", [32, 64, 128])
@pytest.mark.parametrize("bits", [4, 8])
def test_flat_path_vs_oracle(native_lib, cuda_device, dt, sym, g, bits):
    """seeded Llama-like slabs through the flat kernel, both arithmetic modes, packed + unpacked"""
    w = datagen.weights((96, 4096), dt, datagen.seed_of("flat", dt, sym, g, bits), std=0.02)
    w[3, :256] = 0.0                                    # all-zero groups
    w[5, 100] = 3.0                                     # outlier
    for arith in ("native", "fp32"):
        qz = mk(bits=bits, group_size=g, symmetric=sym, arith=arith)
        got = qz.quantize(w, pack=True)
        want = O.pack_result(O.group_quant_vec(w, bits, g, sym, True, arith=arith))
        assert_quant_equal(got, want, f"{dt}/{arith}", keys=("tensor_q", "scales", "zero_points", "qweight", "qzeros"))


@pytest.mark.parametrize("shape", [(7, 300), (5, 1000), (3, 8 * 128 + 64), (1, 129)])
def test_generic_path_packed_vs_oracle(native_lib, cuda_device, shape):
    for dt in ("bf16", "fp32"):
        for sym in (False, True):
            w = datagen.weights(shape, dt, datagen.seed_of("gen", shape, dt, sym), offset=0.1)
            got = mk(symmetric=sym).quantize(w, pack=True)
            want = O.pack_result(O.group_quant_vec(w, 4, 128, sym, True))
            assert_quant_equal(got, want, f"{shape}/{dt}", keys=("tensor_q", "scales", "zero_points", "qweight", "qzeros"))


def test_magnitude_sweep_vs_oracle(native_lib, cuda_device):
    """scales from denormal-ish to huge: exercises the exact (non-hoisted) division path switch"""
    for dt in ("bf16", "fp16", "fp32"):
        for e in (-30, -20, -12, -6, 0, 4):
            if dt == "fp16" and e < -20:
                continue
            w = datagen.weights((16, 1024), dt, datagen.seed_of("mag", dt, e), std=10.0 ** e)
            for sym in (False, True):
                got = mk(symmetric=sym).quantize(w)
                assert_quant_equal(got, O.group_quant_vec(w, 4, 128, sym, True), f"{dt}/1e{e}/{sym}")


def test_offset_groups_vs_oracle(native_lib, cuda_device):
    """all-positive / all-negative / tiny-range groups (zero point clamps, scale floor)"""
    base = datagen.weights((32, 2048), "fp32", 99, std=0.02)
    for off, mul in ((1.0, 1.0), (-1.0, 1.0), (0.5, 1e-6), (100.0, 1e-3), (0.0, 0.0)):
        for dt in ("bf16", "fp16", "fp32"):
            w = (base * mul + off).to(datagen.DTYPES[dt])
            got = mk(symmetric=False).quantize(w)
            assert_quant_equal(got, O.group_quant_vec(w, 4, 128, False, True), f"{dt}/{off}/{mul}")


def test_full_microbench_shape_properties(native_lib, cuda_device):
    """8192 x 28672 (BASELINE.json configs[4]): row-slice parity vs the oracle, pack/unpack identity,
    shard invariance (quantizing row blocks separately gives the same bytes), de-quantization bound."""
    dev = cuda_device
    C, K, g = 8192, 28672, 128
    gen = torch.Generator(device=dev).manual_seed(1234)
    w = (torch.randn((C, K), generator=gen, device=dev, dtype=torch.float32) * 0.02).to(torch.bfloat16)
    for arith in ("native", "fp32"):
        qz = mk(symmetric=False, arith=arith)
        full = qz._quantize_device(w, pack=True)
        rows = torch.tensor([0, 1, 17, 4095, 4096, 8191] + list(range(100, 164)), device=dev)
        sub = w[rows].cpu()
        want = O.pack_result(O.group_quant_vec(sub, 4, g, False, True, arith=arith))
        for k in ("tensor_q", "scales", "zero_points", "qweight", "qzeros"):
            assert_same(full[k][rows].cpu(), want[k], f"{arith}/{k}")
        # pack / unpack identity over the whole tensor, on device
        sh = torch.arange(8, device=dev, dtype=torch.int32) * 4
        unp = ((full["qweight"].unsqueeze(-1) >> sh) & 0xF).reshape(C, K)
        assert torch.equal(unp, full["tensor_q"])
        unz = ((full["qzeros"].unsqueeze(-1) >> sh) & 0xF).reshape(C, -1)[:, : K // g]
        assert torch.equal(unz, full["zero_points"])
        del unp, unz
        # shard invariance: 4 row blocks quantized independently == the full result
        for b in range(4):
            sl = slice(b * 2048, (b + 1) * 2048)
            part = qz._quantize_device(w[sl].contiguous(), pack=True, unpacked=False)
            assert torch.equal(part["qweight"], full["qweight"][sl])
            assert torch.equal(part["qzeros"], full["qzeros"][sl])
            assert torch.equal(part["scales"], full["scales"][sl])
        # codes in range; reconstruction bound: the zero point is rounded (grid shifts by <= s/2) and
        # edge values are clamped, so |w - (q - zp) * s| <= s (+ slack for bf16-rounded arithmetic)
        q = full["tensor_q"]
        assert int(q.min()) >= 0 and int(q.max()) <= 15
        s = full["scales"].float().repeat_interleave(g, 1)
        z = full["zero_points"].float().repeat_interleave(g, 1)
        err = (w.float() - (q.float() - z) * s).abs()
        slack = 1.02 if arith == "fp32" else 1.15
        assert bool((err <= s * slack + 1e-6).all())
        assert float((err / s).mean()) < 0.3                     # typical error is ~ s/4
        del full, q, s, z, err
    torch.cuda.empty_cache()


def test_bf16_to_fp16_exhaustive_and_bulk(native_lib, cuda_device):
    from awq_quantizer.utils.tensor_utils import convert_bf16_to_fp16
    gold = np.load(os.path.join(GOLD, "convert.npz"))["fp16_bits"]
    allbits = torch.arange(65536, dtype=torch.int32).to(torch.int16).view(torch.bfloat16)
    got = convert_bf16_to_fp16(allbits)
    assert got.dtype == torch.float16 and got.device.type == "cpu"
    assert_same(got, torch.from_numpy(gold).view(torch.float16), "bf16->fp16 exhaustive")
    f32 = torch.ones(4)
    assert convert_bf16_to_fp16(f32) is f32             # identity on non-bf16 (tensor_utils.py:20-22)
    # ragged / unaligned lengths, device in -> device out
    for n in (1, 7, 8, 9, 8191, 8192 * 3 + 5):
        x = datagen.weights((n,), "bf16", n, std=300.0).to(cuda_device)
        y = convert_bf16_to_fp16(x)
        assert y.device == x.device
        assert_same(y.cpu(), x.cpu().to(torch.float16), f"n={n}")
        y2 = convert_bf16_to_fp16(x[1:]) if n > 1 else None      # 2-byte-aligned base -> scalar kernel
        if y2 is not None:
            assert_same(y2.cpu(), x[1:].cpu().to(torch.float16), f"n={n} unaligned")


def test_reentrant_from_thread_pool(native_lib, cuda_device):
    """main.py:609-621 calls quantize() from ThreadPoolExecutor workers sharing one quantizer"""
    qz = mk(symmetric=False)
    ws = [datagen.weights((64, 1024 + 128 * i), "bf16", 500 + i) for i in range(12)]
    with cf.ThreadPoolExecutor(max_workers=4) as ex:
        outs = list(ex.map(qz.quantize, ws))
    for w, got in zip(ws, outs):
        assert_quant_equal(got, O.group_quant_vec(w, 4, 128, False, True))


def test_quantize_model_and_input_not_mutated(native_lib, cuda_device):
    qz = mk(symmetric=True)
    w = datagen.weights((8, 256), "bf16", 1)
    keep = w.clone()
    out = qz.quantize_model({"a": w, "skip": torch.zeros(3, dtype=torch.int64), "b": w.float()})
    assert sorted(out) == ["a", "b"]
    assert torch.equal(w.view(torch.int16), keep.view(torch.int16))
    assert_quant_equal(out["a"], O.group_quant_vec(w, 4, 128, True, True))
    nc = datagen.weights((256, 6), "bf16", 2).t()                 # non-contiguous input
    assert_quant_equal(qz.quantize(nc), O.group_quant_vec(nc.contiguous(), 4, 128, True, True))
    with pytest.raises(RuntimeError):
        qz.quantize(torch.zeros(0, 5))


def test_fp64_input(native_lib, cuda_device):
    w = datagen.weights((4, 300), "fp64", 77, offset=0.1)
    for sym in (False, True):
        assert_quant_equal(mk(symmetric=sym).quantize(w), O.group_quant_vec(w, 4, 128, sym, True))


def test_c_abi_direct_device_pointers(native_lib, cuda_device):
    """call the C ABI the way a non-Python host would: raw device pointers + stream"""
    from awq_quantizer import _native as N
    dev = cuda_device
    w = datagen.weights((32, 1024), "bf16", 5).to(dev)
    scales = torch.empty((32, 8), dtype=torch.float16, device=dev)
    zp = torch.empty((32, 8), dtype=torch.int32, device=dev)
    qw = torch.empty((32, 128), dtype=torch.int32, device=dev)
    qz = torch.empty((32, 1), dtype=torch.int32, device=dev)
    st = torch.cuda.Stream(dev)
    st.wait_stream(torch.cuda.current_stream(dev))
    rc = native_lib.awqk_group_quant(w.data_ptr(), N.BF16, 32, 1024, 128, 4, 0, N.ARITH_NATIVE, None,
                                     qw.data_ptr(), scales.data_ptr(), zp.data_ptr(), qz.data_ptr(), None,
                                     st.cuda_stream)
    assert rc == 0
    st.synchronize()
    want = O.pack_result(O.group_quant_vec(w.cpu(), 4, 128, False, True))
    assert_same(qw.cpu(), want["qweight"]); assert_same(qz.cpu(), want["qzeros"])
    assert_same(scales.cpu(), want["scales"]); assert_same(zp.cpu(), want["zero_points"])
    # packed de-quantization == unpacked de-quantization == oracle
    out = torch.empty((32, 1024), dtype=torch.float32, device=dev)
    assert native_lib.awqk_dequant_packed(qw.data_ptr(), scales.data_ptr(), qz.data_ptr(), 32, 1024, 128, 4, 0,
                                          out.data_ptr(), None) == 0
    torch.cuda.synchronize()
    assert_same(out.cpu(), O.dequant_vec({**want, "group_size": torch.tensor(128)}))
    # host pointer where a device pointer is expected -> error code, not a crash
    host = torch.zeros(4, 128, dtype=torch.bfloat16)
    assert native_lib.awqk_group_quant(host.data_ptr(), N.BF16, 4, 128, 128, 4, 0, 0, None, None,
                                       scales.data_ptr(), None, None, None, None) < 0


def test_model_arena_pipeline_vs_oracle(native_lib, cuda_device):
    """quantize_model(pack=True): flat arena + chunked H2D/K1/D2H pipeline (awqk_pipe_quant_host) for
    tensors whose rows are whole groups / whole zero words, per-tensor path for the rest"""
    from awq_quantizer.quantization.arena import HostArena
    shapes = {"a.weight": (256, 1024), "b.weight": (64, 4096), "bias": (1024,), "odd_g4": (96, 512),
              "ragged": (7, 300), "conv": (32, 2, 1024), "tiny": (10, 10)}
    tensors = {n: datagen.weights(s, "bf16", datagen.seed_of("arena", n)) for n, s in shapes.items()}
    tensors["f32.weight"] = datagen.weights((16, 2048), "fp32", 9)
    tensors["ints"] = torch.zeros(8, 128, dtype=torch.int32)          # logged and skipped
    for sym, arith in ((False, "native"), (True, "native"), (False, "fp32")):
        qz = mk(symmetric=sym, arith=arith)
        # small chunk size -> several chunks per arena, exercising slot reuse in the pipeline
        out = qz.quantize_model(tensors, pack=True, chunk_bytes=1 << 16)
        assert sorted(out) == sorted(n for n in tensors if n != "ints")
        for n, t in tensors.items():
            if n == "ints":
                continue
            if t.numel() < 128:
                want = O.group_quant_vec(t, 4, 128, sym, True, arith=arith)
                assert_same(out[n]["scales"], want["scales"], n)
                continue
            want = O.pack_result(O.group_quant_vec(t, 4, 128, sym, True, arith=arith))
            for k in ("qweight", "qzeros", "scales"):
                assert_same(out[n][k], want[k], f"{n}/{k}/{sym}/{arith}")
            assert "tensor_q" not in out[n]
    # zero-copy form: tensors already live in a HostArena
    eligible = {n: t for n, t in tensors.items() if n in ("a.weight", "b.weight", "bias", "conv")}
    arena = HostArena.from_tensors(eligible)
    out = mk(symmetric=False).quantize_model(arena, pack=True)
    for n, t in eligible.items():
        want = O.pack_result(O.group_quant_vec(t, 4, 128, False, True))
        for k in ("qweight", "qzeros", "scales"):
            assert_same(out[n][k], want[k], f"arena/{n}/{k}")


def test_pipe_c_abi_unpacked_and_errors(native_lib, cuda_device):
    import ctypes as C
    from awq_quantizer import _native as N
    w = datagen.weights((64, 2048), "bf16", 31).pin_memory()
    q = torch.empty((64, 2048), dtype=torch.int32).pin_memory()
    qp = torch.empty((64, 256), dtype=torch.int32).pin_memory()
    sc = torch.empty((64, 16), dtype=torch.float16).pin_memory()
    zp = torch.empty((64, 16), dtype=torch.int32).pin_memory()
    zq = torch.empty((64, 2), dtype=torch.int32).pin_memory()
    h = C.c_void_p()
    assert native_lib.awqk_pipe_create(0, 1 << 16, C.byref(h)) == 0
    try:
        rc = native_lib.awqk_pipe_quant_host(h, w.data_ptr(), N.BF16, 64, 2048, 128, 4, 0, N.ARITH_NATIVE,
                                             q.data_ptr(), qp.data_ptr(), sc.data_ptr(), zp.data_ptr(), zq.data_ptr())
        assert rc == 0 and native_lib.awqk_pipe_sync(h) == 0
        want = O.pack_result(O.group_quant_vec(w, 4, 128, False, True))
        assert_same(q, want["tensor_q"]); assert_same(qp, want["qweight"]); assert_same(sc, want["scales"])
        assert_same(zp, want["zero_points"]); assert_same(zq, want["qzeros"])
        # ragged rows are not a flat layout -> explicit error code, never a silent fallback
        assert native_lib.awqk_pipe_quant_host(h, w.data_ptr(), N.BF16, 64, 2047, 128, 4, 0, 0, None, qp.data_ptr(),
                                               sc.data_ptr(), None, None) == -4
    finally:
        native_lib.awqk_pipe_destroy(h)


@pytest.mark.parametrize("bits,g", [(4, 32), (4, 64), (8, 128), (8, 32)])
def test_model_arena_other_bits_and_groups(native_lib, cuda_device, bits, g):
    shapes = {"a": (64, 2048), "b": (16, 4096), "bias": (2048,), "rowmode": (48, 3 * g)}
    tensors = {n: datagen.weights(s, "bf16", datagen.seed_of("arena2", n, bits, g)) for n, s in shapes.items()}
    tensors["h"] = datagen.weights((32, 1024), "fp16", 5)
    qz = mk(bits=bits, group_size=g, symmetric=False)
    out = qz.quantize_model(tensors, pack=True, chunk_bytes=1 << 16)
    for n, t in tensors.items():
        want = O.pack_result(O.group_quant_vec(t, bits, g, False, True))
        for k in ("qweight", "qzeros", "scales"):
            assert_same(out[n][k], want[k], f"{n}/{k}/b{bits}/g{g}")


def test_dequant_paths(native_lib, cuda_device):
    """row-structured fast path (K % 4 == 0, g % 4 == 0) and the generic element path"""
    for shape, g in (((64, 1024), 128), ((3, 4100), 100), ((5, 1030), 128), ((2, 3, 512), 64), ((7, 129), 7)):
        w = datagen.weights(shape, "bf16", datagen.seed_of("dq", shape, g), offset=0.05)
        qz = mk(group_size=g, symmetric=False)
        r = qz.quantize(w)
        assert_same(qz.dequantize(r), O.dequant_vec(r), f"{shape}/g{g}")


def test_more_than_2_31_elements(native_lib, cuda_device):
    """64-bit indexing: a 2.7e9-element bf16 tensor through the flat TMA kernel (row-slice parity vs the
    oracle, pack/unpack identity by blocks)"""
    dev = cuda_device
    C, K, g = 40960, 65536, 128
    assert C * K > 2 ** 31
    gen = torch.Generator(device=dev).manual_seed(7)
    w = torch.empty((C, K), dtype=torch.bfloat16, device=dev)
    for r0 in range(0, C, 4096):                                   # generate in blocks (fp32 scratch stays small)
        w[r0:r0 + 4096] = (torch.randn((4096, K), generator=gen, device=dev) * 0.02).to(torch.bfloat16)
    qz = mk(symmetric=False)
    full = qz._quantize_device(w, pack=True, unpacked=False)
    rows = torch.tensor([0, 1, 32767, 32768, 32769, 40959], device=dev)    # around the 2^31-element boundary
    sub = w[rows].cpu()
    want = O.pack_result(O.group_quant_vec(sub, 4, g, False, True))
    for k in ("scales", "zero_points", "qweight", "qzeros"):
        assert_same(full[k][rows].cpu(), want[k], k)
    # a second, independent quantization of the upper half must equal the slices of the full result
    part = qz._quantize_device(w[32768:].contiguous(), pack=True, unpacked=False)
    assert torch.equal(part["qweight"], full["qweight"][32768:])
    assert torch.equal(part["scales"], full["scales"][32768:]) and torch.equal(part["qzeros"], full["qzeros"][32768:])
    del w, full, part
    torch.cuda.empty_cache()


def _guarded(nbytes, dev):
    """(buffer, view): `view` is nbytes in the middle of a buffer whose 4 KiB borders hold a canary"""
    pad = 4096
    buf = torch.full((nbytes + 2 * pad,), 0xA5, dtype=torch.uint8, device=dev)
    return buf, buf[pad:pad + nbytes]


def _canaries_intact(buf, nbytes):
    pad = 4096
    return bool((buf[:pad] == 0xA5).all()) and bool((buf[pad + nbytes:] == 0xA5).all())


def test_no_out_of_bounds_writes(native_lib, cuda_device):
    """compute-sanitizer is closed on this pool: guard every output buffer with canaries instead
    (partial CTA tiles, row mode, generic path, convert, de-quantizer, fake-quant delta)"""
    from awq_quantizer import _native as N
    dev = cuda_device
    L = native_lib
    for (C, K), g, bits in (((72, 1152), 128, 4), ((9, 1024), 32, 4), ((5, 640), 64, 8), ((7, 300), 128, 4), ((1, 128), 128, 4)):
        w = datagen.weights((C, K), "bf16", datagen.seed_of("oob", C, K)).to(dev)
        G = -(-K // g)
        per = 32 // bits
        sizes = {"q": C * K * 4, "qp": C * (-(-K // per)) * 4, "s": C * G * 2, "z": C * G * 4, "zq": C * (-(-G // per)) * 4}
        bufs = {k: _guarded(v, dev) for k, v in sizes.items()}
        for unpacked in (True, False):
            rc = L.awqk_group_quant(w.data_ptr(), N.BF16, C, K, g, bits, 0, N.ARITH_NATIVE,
                                    bufs["q"][1].data_ptr() if unpacked else None, bufs["qp"][1].data_ptr(),
                                    bufs["s"][1].data_ptr(), bufs["z"][1].data_ptr(), bufs["zq"][1].data_ptr(), None, None)
            assert rc == 0
            torch.cuda.synchronize()
            for k, (b, _) in bufs.items():
                assert _canaries_intact(b, sizes[k]), (C, K, g, bits, unpacked, k)
        want = O.pack_result(O.group_quant_vec(w.cpu(), bits, g, False, True))
        got_qp = bufs["qp"][1].view(torch.int32).reshape(C, -1).cpu()
        assert_same(got_qp, want["qweight"], f"{C}x{K}")
        # dequant + convert
        out_b, out_v = _guarded(C * K * 4, dev)
        assert L.awqk_dequant(bufs["q"][1].data_ptr(), bufs["s"][1].data_ptr(), bufs["z"][1].data_ptr(), C, K, g,
                              out_v.data_ptr(), None) == 0
        h_b, h_v = _guarded(C * K * 2, dev)
        assert L.awqk_bf16_to_fp16(w.data_ptr(), h_v.data_ptr(), C * K, None) == 0
        torch.cuda.synchronize()
        assert _canaries_intact(out_b, C * K * 4) and _canaries_intact(h_b, C * K * 2)
    # fake-quant delta and the GEMM's err vector
    C, K, n = 37, 1280, 3
    w = datagen.weights((C, K), "bf16", 3).to(dev)
    s = (torch.rand((n, K), device=dev) + 0.5)
    dw_b, dw_v = _guarded(n * C * K * 2, dev)
    assert L.awqk_fakequant_delta(w.data_ptr(), N.BF16, C, K, 128, 4, 0, s.data_ptr(), n, dw_v.data_ptr(), None) == 0
    torch.cuda.synchronize()
    assert _canaries_intact(dw_b, n * C * K * 2)
    x = datagen.activations(70, K, "bf16", 5).to(dev)
    e_b, e_v = _guarded(n * 8, dev)
    e_v.zero_()
    assert L.awqk_sqerr_gemm(x.data_ptr(), dw_v.data_ptr(), 70, C, K, n, e_v.data_ptr(), None) == 0
    torch.cuda.synchronize()
    assert _canaries_intact(e_b, n * 8) and float(e_v.view(torch.float64).min()) > 0
    # the one-call search: minimum workspace (ring of delta slabs + counters) and all outputs between canaries
    import ctypes
    for (C, K, T) in ((37, 1280, 70), (300, 1024, 260)):
        w = datagen.weights((C, K), "bf16", 3).to(dev)
        x = datagen.activations(T, K, "bf16", 5).to(dev)
        mn = ctypes.c_size_t(0)
        L.awqk_workspace_bytes(C, K, T, n, 0, ctypes.byref(mn))
        G = K // 128
        sizes = {"ws": mn.value, "err": n * 8, "best": 4, "s": K * 4, "qp": C * K // 8 * 4, "sc": C * G * 2, "z": C * G * 4,
                 "zq": C * (-(-G // 8)) * 4}
        bufs = {k: _guarded(v, dev) for k, v in sizes.items()}
        v = {k: b[1] for k, b in bufs.items()}
        assert v["ws"].data_ptr() % 256 == 0
        rc = L.awqk_scale_search(w.data_ptr(), N.BF16, C, K, x.data_ptr(), T, None, n, 128, 4, 0, v["err"].data_ptr(),
                                 v["best"].data_ptr(), v["s"].data_ptr(), None, v["qp"].data_ptr(), v["sc"].data_ptr(),
                                 v["z"].data_ptr(), v["zq"].data_ptr(), v["ws"].data_ptr(), mn.value, None)
        assert rc == 0, rc
        torch.cuda.synchronize()
        for k, (b, _) in bufs.items():
            assert _canaries_intact(b, sizes[k]), (C, K, T, k)


@pytest.mark.parametrize("shape", [(64, 256), (136, 1024), (8, 8 * 128), (200, 520)])
def test_autoawq_export_vs_oracle(native_lib, cuda_device, shape):
    from awq_quantizer.quantization.export import to_autoawq_gemm
    w = datagen.weights(shape, "bf16", datagen.seed_of("awq", shape))
    for sym in (False, True):
        r = mk(symmetric=sym).quantize(w, pack=True)
        got = to_autoawq_gemm(r, device="cuda:0")
        want = O.to_autoawq_gemm(O.group_quant_vec(w, 4, 128, sym, True))
        for k in ("qweight", "qzeros", "scales"):
            assert_same(got[k], want[k], f"{shape}/{sym}/{k}")
    with pytest.raises(ValueError):
        to_autoawq_gemm(mk(bits=8).quantize(w, pack=True))


def test_default_quantize_model_is_pipelined_and_reference_exact(native_lib, cuda_device):
    """quantize_model(tensors) with unchanged arguments: reference layout, bit-exact, input order kept,
    bad tensors skipped -- now through the arena pipeline for every whole-group tensor"""
    shapes = {"w1": (64, 1024), "odd_g": (96, 512), "bias": (1024,), "ragged": (7, 300), "tiny": (10, 10), "conv": (16, 3, 256)}
    tensors = {n: datagen.weights(s, "bf16", datagen.seed_of("dflt", n)) for n, s in shapes.items()}
    tensors["bad"] = torch.zeros(4, 4, dtype=torch.int64)
    tensors["h"] = datagen.weights((8, 256), "fp16", 3)
    for sym in (False, True):
        qz = mk(symmetric=sym)
        out = qz.quantize_model(tensors, chunk_bytes=1 << 16)
        assert list(out) == [n for n in tensors if n != "bad"]
        for n, r in out.items():
            want = O.group_quant_vec(tensors[n], 4, 128, sym, True)
            assert_quant_equal(r, want, n)
            assert r["tensor_q"].shape == tensors[n].shape and r["tensor_q"].device.type == "cpu"
            assert int(r["bits"]) == 4 and bool(r["symmetric"]) == sym
            if r["scales"].dim() == 2:
                assert_same(qz.dequantize(r), O.dequant_vec(want), n + "/dequant")
        loop = qz.quantize_model(tensors, pipeline=False)
        for n in out:
            assert_quant_equal(out[n], loop[n], n + "/loop")


@pytest.mark.parametrize("chunk_bytes", [1 << 16, 3 << 16, 32 << 20])
@pytest.mark.parametrize("layout", ["packed", "reference", "both"])
def test_gather_pipeline_from_pageable_tensors(native_lib, cuda_device, chunk_bytes, layout):
    """awqk_pipe_quant_gather: virtual arena over pageable tensors, pinned bounce rings inside the pipe, results
    drained into pageable arrays -- chunks ending inside tensors, at tensor ends and inside tile padding; many
    more chunks than ring slots"""
    from awq_quantizer.quantization.arena import HostArena, quantize_arena
    shapes = {"a": (24, 1024), "b": (1024,), "c": (200, 2048), "d": (8, 3, 1024), "e": (3, 1024), "h": (16, 1024),
              "f": (1, 9 * 8192 + 1024)}
    tensors = {n: datagen.weights(s, "fp16" if n == "h" else "bf16", datagen.seed_of("gather", n)) for n, s in shapes.items()}
    arena = HostArena.for_tensors(tensors)
    assert not arena.buffers                                   # layout only: nothing was pinned
    packed, unpacked = layout in ("packed", "both"), layout in ("reference", "both")
    for rep in range(2):                                       # second call reuses the rings
        res = quantize_arena(arena, bits=4, group_size=128, symmetric=bool(rep), arith="native", device=cuda_device,
                             chunk_bytes=chunk_bytes, packed=packed, unpacked=unpacked, want_zero_points=True,
                             sources=tensors, pin_results=bool(rep))      # rep 1: direct D2H into pinned results
        for n, t in tensors.items():
            want = O.pack_result(O.group_quant_vec(t, 4, 128, bool(rep), True))
            keys = ("scales", "zero_points") + (("qweight", "qzeros") if packed else ()) + (("tensor_q",) if unpacked else ())
            assert_quant_equal(res[n], want, f"{n}/{rep}", keys=keys)
            assert res[n]["scales"].is_pinned() == bool(rep)


def test_gather_c_abi_errors(native_lib, cuda_device):
    import ctypes as C
    from awq_quantizer import _native as N
    w = datagen.weights((4, 1000), "bf16", 1)                   # 4000 elements: not a multiple of the group size
    sc = torch.empty(64, dtype=torch.float16)
    h = C.c_void_p()
    assert native_lib.awqk_pipe_create(0, 1 << 16, C.byref(h)) == 0
    try:
        ptrs = (C.c_void_p * 1)(w.data_ptr())
        nums = (C.c_int64 * 1)(4000)
        args = (0, N.BF16, 128, 4, 0, N.ARITH_NATIVE, None, None, sc.data_ptr(), None, None)
        assert native_lib.awqk_pipe_quant_gather(h, 1, ptrs, nums, *args) == -1
        nums[0] = 0
        assert native_lib.awqk_pipe_quant_gather(h, 1, ptrs, nums, *args) == -1
        assert native_lib.awqk_pipe_quant_gather(h, 0, ptrs, nums, *args) == -1
        assert native_lib.awqk_pipe_quant_gather(None, 1, ptrs, nums, *args) == -1
        nums[0] = 3840                                          # rows of 384 = 3 groups: not a short-row class
        assert native_lib.awqk_pipe_quant_gather(h, 1, ptrs, nums, 384, *args[1:]) == -1
        assert native_lib.awqk_pipe_quant_gather(h, 1, ptrs, nums, 1024, *args[1:]) == -1     # 8 groups = a whole word
    finally:
        native_lib.awqk_pipe_destroy(h)


@pytest.mark.parametrize("dt", ["bf16", "fp16"])
def test_rows_of_1_2_4_groups_pack_zero_words_in_kernel(native_lib, cuda_device, dt):
    """K = 1, 2 or 4 groups: K1 v2 writes the row-padded qzeros word itself (no int32 zero points needed, no
    second kernel); compared with the oracle's per-row packing"""
    from awq_quantizer import _native as N
    for (C, K), g in (((72, 512), 128), ((9, 256), 128), ((33, 128), 128), ((40, 128), 32), ((8, 64), 32), ((5, 32), 32),
                      ((130, 256), 64)):
        for sym in (False, True):
            w = datagen.weights((C, K), dt, datagen.seed_of("rowzp", C, K, g), offset=0.01)
            G = K // g
            wd = w.to(cuda_device)
            qp = torch.full((C, K // 8), 5, dtype=torch.int32, device=cuda_device)
            sc = torch.zeros((C, G), dtype=torch.float16, device=cuda_device)
            zq = torch.full((C, 1), 5, dtype=torch.int32, device=cuda_device)
            rc = native_lib.awqk_group_quant(wd.data_ptr(), N.dtype_code(w.dtype), C, K, g, 4, int(sym), N.ARITH_NATIVE,
                                             None, qp.data_ptr(), sc.data_ptr(), None, zq.data_ptr(), None, None)
            assert rc == 0, (rc, C, K, g)
            torch.cuda.synchronize()
            want = O.pack_result(O.group_quant_vec(w, 4, g, sym, True))
            assert_same(qp.cpu(), want["qweight"], f"{C}x{K}/g{g}/qweight")
            assert_same(sc.cpu(), want["scales"], f"{C}x{K}/g{g}/scales")
            assert_same(zq.cpu(), want["qzeros"], f"{C}x{K}/g{g}/qzeros")


@pytest.mark.parametrize("dt", ["bf16", "fp16"])
@pytest.mark.parametrize("arith", ["native", "fp32"])
@pytest.mark.parametrize("sym", [False, True])
def test_int8_through_tma_kernel(native_lib, cuda_device, dt, arith, sym):
    """bits = 8 on the TMA path: packed words of 4 codes, 4 zero points per word (1- and 2-group rows packed in the
    kernel), the reference's int32 codes, special values, magnitudes that push groups onto the exact path"""
    from awq_quantizer import _native as N
    ar = N.ARITH_FP32 if arith == "fp32" else N.ARITH_NATIVE
    for (C, K), g, scale in (((40, 1024), 128, 0.02), ((9, 256), 128, 1.0), ((33, 128), 128, 300.0), ((16, 64), 32, 0.02),
                             ((130, 512), 64, 1e-3), ((7, 4096), 32, 5.0)):
        w = datagen.weights((C, K), "fp32", datagen.seed_of("i8", C, K, g), std=scale, offset=0.1 * scale)
        w[0, :g] = 0.0
        w[1, 0] = float("nan")
        w[2, 1] = float("inf")
        w[3, :g] = 0.75 * scale
        w = w.to(datagen.DTYPES[dt])
        G = K // g
        wd = w.to(cuda_device)
        q = torch.full((C, K), 9, dtype=torch.int32, device=cuda_device)
        qp = torch.full((C, K // 4), 9, dtype=torch.int32, device=cuda_device)
        sc = torch.zeros((C, G), dtype=torch.float16, device=cuda_device)
        zp = torch.full((C, G), 9, dtype=torch.int32, device=cuda_device)
        zq = torch.full((C, -(-G // 4)), 9, dtype=torch.int32, device=cuda_device)
        needs_zp = G % 4 != 0 and G not in (1, 2)
        for unpacked in (False, True):
            rc = native_lib.awqk_group_quant(wd.data_ptr(), N.dtype_code(w.dtype), C, K, g, 8, int(sym), ar,
                                             q.data_ptr() if unpacked else None, qp.data_ptr(), sc.data_ptr(),
                                             zp.data_ptr() if (unpacked or needs_zp) else None, zq.data_ptr(), None, None)
            assert rc == 0, (rc, C, K, g)
            torch.cuda.synchronize()
            want = O.pack_result(O.group_quant_vec(w, 8, g, sym, True, arith=arith))
            what = f"{C}x{K}/g{g}/{unpacked}"
            assert_same(qp.cpu(), want["qweight"], what + "/qweight")
            assert_same(sc.cpu(), want["scales"], what + "/scales")
            assert_same(zq.cpu(), want["qzeros"], what + "/qzeros")
            if unpacked:
                assert_same(q.cpu(), want["tensor_q"], what + "/tensor_q")
                assert_same(zp.cpu(), want["zero_points"].reshape(C, G), what + "/zero_points")


@pytest.mark.parametrize("sym", [False, True])
@pytest.mark.parametrize("g", [32, 64, 128])
def test_fp32_input_through_tma_kernel(native_lib, cuda_device, sym, g):
    """fp32 weights on the TMA path (8 LDS.128 per thread, XOR-swizzled chunk order, halves of a packed word swapped on
    odd lanes): packed words, int32 codes, zero points, special values, partial tiles, many tiles per CTA"""
    from awq_quantizer import _native as N
    for (C, K), scale in (((40, 1024), 0.02), ((3, 128 * 7), 1.0), ((600, 8192), 0.02), ((33, 2048), 1e-20), ((9, 512), 3e4)):
        w = datagen.weights((C, K), "fp32", datagen.seed_of("f32tma", C, K, g), std=scale, offset=0.3 * scale)
        w[0, :g] = 0.0
        w[1, 5] = float("nan")
        w[2, 77] = float("-inf")
        w[-1, -g:] = 1.25 * scale
        G = K // g
        wd = w.to(cuda_device)
        q = torch.full((C, K), 9, dtype=torch.int32, device=cuda_device)
        qp = torch.full((C, K // 8), 9, dtype=torch.int32, device=cuda_device)
        sc = torch.zeros((C, G), dtype=torch.float16, device=cuda_device)
        zp = torch.full((C, G), 9, dtype=torch.int32, device=cuda_device)
        flat_zq = G % 8 == 0 or G in (1, 2, 4)
        zq = torch.full((C, -(-G // 8)), 9, dtype=torch.int32, device=cuda_device)
        want = O.pack_result(O.group_quant_vec(w, 4, g, sym, True))
        for unpacked in (False, True):
            rc = native_lib.awqk_group_quant(wd.data_ptr(), N.FP32, C, K, g, 4, int(sym), N.ARITH_NATIVE,
                                             q.data_ptr() if unpacked else None, qp.data_ptr(), sc.data_ptr(),
                                             zp.data_ptr() if (unpacked or not flat_zq) else None, zq.data_ptr(), None, None)
            assert rc == 0, (rc, C, K, g)
            torch.cuda.synchronize()
            what = f"{C}x{K}/g{g}/{unpacked}"
            assert_same(qp.cpu(), want["qweight"], what + "/qweight")
            assert_same(sc.cpu(), want["scales"], what + "/scales")
            assert_same(zq.cpu(), want["qzeros"], what + "/qzeros")
            if unpacked:
                assert_same(q.cpu(), want["tensor_q"], what + "/tensor_q")
                assert_same(zp.cpu(), want["zero_points"].reshape(C, G), what + "/zero_points")


def test_quantize_model_from_worker_threads(native_lib, cuda_device):
    """the CLI's --multi_gpu mode calls quantize_model from pool threads (main.py:596-621): every thread owns its pipe
    (streams, device slots, pinned rings), which is destroyed when the thread ends"""
    import gc
    from concurrent.futures import ThreadPoolExecutor
    tensors = {f"t{i}": datagen.weights((32 + 8 * i, 1024), "bf16", 500 + i) for i in range(6)}
    want = {n: O.pack_result(O.group_quant_vec(t, 4, 128, False, True)) for n, t in tensors.items()}

    def work(k):
        qz = mk(symmetric=False)
        out = qz.quantize_model(tensors, pack=True, chunk_bytes=1 << 16, keep_unpacked=bool(k & 1))
        for n in tensors:
            assert_quant_equal(out[n], want[n], f"{k}/{n}", keys=("scales", "qweight", "qzeros") + (("tensor_q",) if k & 1 else ()))
        return k

    for _ in range(2):                      # the second pool runs on fresh threads: fresh pipes, old ones released
        with ThreadPoolExecutor(max_workers=4) as ex:
            assert sorted(ex.map(work, range(8))) == list(range(8))
        gc.collect()


@pytest.mark.parametrize("keep", [False, True])
def test_short_row_tensors_through_gather(native_lib, cuda_device, keep):
    """rows of 1 / 2 / 4 groups (OPT's K = 512 tensors) join the gather pipeline in a class of their own: qzeros is one
    zero-padded word per row, written by K1"""
    tensors = {"embed": datagen.weights((300, 512), "bf16", 1), "proj": datagen.weights((64, 512), "bf16", 2),
               "k256": datagen.weights((40, 256), "bf16", 3), "k128": datagen.weights((9, 128), "fp16", 4),
               "full": datagen.weights((16, 1024), "bf16", 5), "odd": datagen.weights((8, 1536), "bf16", 6)}
    for sym in (False, True):
        out = mk(symmetric=sym).quantize_model(tensors, pack=True, chunk_bytes=1 << 16, keep_unpacked=keep)
        assert list(out) != [] and sorted(out) == sorted(tensors)
        for n, t in tensors.items():
            want = O.pack_result(O.group_quant_vec(t, 4, 128, sym, True))
            keys = ("scales", "qweight", "qzeros") + (("tensor_q", "zero_points") if keep else ())
            assert_quant_equal(out[n], want, f"{n}/{sym}", keys=keys)


def test_quantize_routes_large_host_tensors_through_the_pipeline(native_lib, cuda_device):
    """AWQQuantizer.quantize(tensor) -- the reference's own per-tensor call -- streams a large host tensor through
    the gather pipeline instead of upload -> kernel -> download; same result dict, bit-exact"""
    cases = [((1100, 1024), "bf16", 4), ((2100, 512), "bf16", 4), ((1 << 20,), "fp16", 4), ((64, 3, 8192), "bf16", 4),
             ((520, 2048), "fp32", 4), ((1100, 1024), "bf16", 8)]
    for shape, dt, bits in cases:
        w = datagen.weights(shape, dt, datagen.seed_of("route", shape, dt, bits))
        for sym in (False, True):
            qz = mk(symmetric=sym, bits=bits)
            assert qz._quantize_host_pipelined(w, cuda_device, False, True) is not None, (shape, dt, bits)
            want = O.pack_result(O.group_quant_vec(w, bits, 128, sym, True))
            ref = qz.quantize(w)                                      # unchanged reference call
            assert sorted(ref) == ["bits", "group_size", "scales", "symmetric", "tensor_q", "zero_points"]
            assert_quant_equal(ref, want, f"{shape}/{dt}/{bits}/{sym}")
            assert ref["tensor_q"].shape == w.shape and ref["tensor_q"].dtype == torch.int32
            assert ref["scales"].shape == want["scales"].shape and ref["zero_points"].dtype == torch.int32
            both = qz.quantize(w, pack=True)
            assert_quant_equal(both, want, "both", keys=("tensor_q", "scales", "zero_points", "qweight", "qzeros"))
            packed = qz.quantize(w, pack=True, keep_unpacked=False)
            assert "tensor_q" not in packed
            assert_quant_equal(packed, want, "packed", keys=("scales", "zero_points", "qweight", "qzeros"))
    small = datagen.weights((64, 1024), "bf16", 3)
    assert mk()._quantize_host_pipelined(small, cuda_device, False, True) is None          # below the threshold
    ragged = datagen.weights((4100, 300), "bf16", 4)
    assert mk()._quantize_host_pipelined(ragged, cuda_device, False, True) is None         # rows are not whole groups
    assert_quant_equal(mk(symmetric=False).quantize(ragged), O.group_quant_vec(ragged, 4, 128, False, True), "ragged")


def test_native_pinned_staging_buffers(native_lib, cuda_device):
    """awqk_host_alloc_pinned: page-locked in place (huge-page mapping + cudaHostRegister), usable for asynchronous
    copies in both directions, 2 MiB aligned, returned to / reused from the process-wide cache, freed on eviction"""
    import ctypes
    from awq_quantizer import _native as N
    nb = (8 << 20) + 4096
    t = N.pinned_take(nb)
    assert t.dtype == torch.uint8 and t.numel() == nb and t.is_pinned()
    assert t.data_ptr() % 4096 == 0                 # (2 MiB on the cudaHostRegister path, a page on the cudaHostAlloc fallback)
    src = torch.arange(nb, dtype=torch.int64).to(torch.uint8)
    t.copy_(src)
    d = torch.empty(nb, dtype=torch.uint8, device=cuda_device)
    d.copy_(t, non_blocking=True)
    back = N.pinned_take(nb)
    assert back.data_ptr() != t.data_ptr()
    back.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    assert torch.equal(back, src)
    ptr = t.data_ptr()
    N.pinned_give_back(t)
    N.pinned_give_back(back)
    again = N.pinned_take(nb)                       # same size: served from the cache
    assert again.data_ptr() in (ptr, back.data_ptr())
    N.pinned_give_back(again)
    N.pinned_give_back(torch.empty(16, dtype=torch.uint8))      # foreign tensors are ignored
    # raw ABI: allocate / free, bad pointers are refused
    out = ctypes.c_void_p()
    assert native_lib.awqk_host_alloc_pinned(1 << 20, ctypes.byref(out)) == 0 and out.value
    assert native_lib.awqk_host_free_pinned(out) == 0
    assert native_lib.awqk_host_alloc_pinned(0, ctypes.byref(out)) == -1
    assert native_lib.awqk_host_free_pinned(None) == 0


@pytest.mark.parametrize("sym", [False, True])
def test_autoawq_export_is_read_correctly_by_vllm(native_lib, cuda_device, sym):
    """The AutoAWQ / vLLM GEMM layout written by quantization/export.py is PINNED against a third-party consumer:
    vLLM's own AWQ de-quantization kernel (vllm._custom_ops.awq_dequantize, the kernel its AWQ linear layers use) must
    reconstruct, from the exported qweight [K, C/8] / qzeros [G, C/8] / scales [G, C], exactly the weight this
    repo's de-quantizer (the reference's arithmetic, K4) reconstructs from the packed K1 result."""
    try:
        from vllm import _custom_ops as vops
    except Exception as e:                                        # pragma: no cover
        pytest.skip(f"vllm is not importable here: {e}")
    from awq_quantizer.quantization import AWQQuantizer
    from awq_quantizer.quantization.export import to_autoawq_gemm
    C, K = 256, 1024
    w = datagen.weights((C, K), "fp16", 77)
    qz = AWQQuantizer(bits=4, group_size=128, symmetric=sym, device="cuda:0", logger_level="ERROR")
    res = qz.quantize(w, pack=True)
    exp = to_autoawq_gemm(res, device=cuda_device)
    try:
        back = vops.awq_dequantize(exp["qweight"].to(cuda_device), exp["scales"].to(cuda_device),
                                   exp["qzeros"].to(cuda_device), 0, 0, 0)
    except Exception as e:                                        # pragma: no cover
        pytest.skip(f"vllm's awq_dequantize is not usable on this box: {e}")
    torch.cuda.synchronize()
    assert back.shape == (K, C) and back.dtype == torch.float16
    mine = qz.dequantize(res)                                     # fp32 [C, K], fp16 multiply like the reference
    got = back.float().t().cpu()
    assert torch.equal(got, mine), float((got - mine).abs().max())
    # and vLLM's AWQ GEMM (the kernel that serves such a checkpoint) computes x . W^T with those weights
    try:
        x = datagen.weights((16, K), "fp16", 5, std=1.0).to(cuda_device)
        y = vops.awq_gemm(x, exp["qweight"].to(cuda_device), exp["scales"].to(cuda_device), exp["qzeros"].to(cuda_device), 8)
    except Exception as e:                                        # pragma: no cover
        pytest.skip(f"vllm's awq_gemm is not usable on this box: {e}")
    torch.cuda.synchronize()
    want = x.float() @ mine.to(cuda_device).t()
    assert y.shape == (16, C)
    err = float((y.float() - want).abs().max() / want.abs().max())
    assert err < 2e-2, err                                        # fp16 accumulation inside the kernel

"""GPU (B200): K1 v2 in column-slab mode (csrc/awqk_group_quant_tma.cu, CS = true) -- the final AWQ pass
`group_quant(fp32(W) * s[k])` -- against the oracle's `quantize_scaled` (the pinned group quantizer
composed with the per-input-channel scale).  Bit-exact for every output, straight through the C ABI."""
import pytest
import torch

from oracle import awq_oracle as O
from tests import datagen
from tests.util import assert_same

pytestmark = pytest.mark.gpu


def _scales(K, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.exp(0.7 * torch.randn(K, generator=g)).to(torch.float32)


def _run(L, N, w, s, g, sym, dev, *, unpacked=True, packed=True, want_zp=True, want_zq=True):
    C, K = w.shape
    G = K // g
    wd, sd = w.to(dev), s.to(dev)
    q = torch.full((C, K), 77, dtype=torch.int32, device=dev) if unpacked else None
    qp = torch.full((C, K // 8), 77, dtype=torch.int32, device=dev) if packed else None
    sc = torch.zeros((C, G), dtype=torch.float16, device=dev)
    zp = torch.full((C, G), 77, dtype=torch.int32, device=dev) if want_zp else None
    zq = torch.full((C, -(-G // 8)), 77, dtype=torch.int32, device=dev) if want_zq else None
    rc = L.awqk_group_quant(wd.data_ptr(), N.dtype_code(w.dtype), C, K, g, 4, int(sym), N.ARITH_FP32,
                            N.ptr(q), N.ptr(qp), sc.data_ptr(), N.ptr(zp), N.ptr(zq), sd.data_ptr(), None)
    assert rc == 0, rc
    torch.cuda.synchronize()
    return {"tensor_q": q, "qweight": qp, "scales": sc, "zero_points": zp, "qzeros": zq}


def _check(got, want, what):
    for k, v in got.items():
        if v is not None:
            assert_same(v.cpu(), want[k], f"{what}/{k}")


@pytest.mark.parametrize("dt", ["bf16", "fp16"])
@pytest.mark.parametrize("sym", [False, True])
@pytest.mark.parametrize("g", [32, 64, 128])
def test_colscale_slab_kernel_bit_exact(native_lib, cuda_device, dt, sym, g):
    from awq_quantizer import _native as N
    for C, K in ((1, 1024), (8, 1024), (13, 2048), (300, 4096), (77, 3072)):
        w = datagen.weights((C, K), dt, datagen.seed_of("cs", C, K, dt))
        s = _scales(K, C + K)
        want = O.pack_result(O.quantize_scaled(w, s, 4, g, sym))
        _check(_run(native_lib, N, w, s, g, sym, cuda_device), want, f"{C}x{K}")
        _check(_run(native_lib, N, w, s, g, sym, cuda_device, unpacked=False), want, f"{C}x{K}/packed only")
        _check(_run(native_lib, N, w, s, g, sym, cuda_device, packed=False, want_zq=False), want, f"{C}x{K}/unpacked only")


@pytest.mark.parametrize("sym", [False, True])
def test_colscale_special_values(native_lib, cuda_device, sym):
    """slow-path groups inside the slab kernel: non-finite values, constant / all-zero groups, huge and tiny
    magnitudes, all-positive rows (clamped zero point), scales spanning 2^-20 .. 2^20"""
    from awq_quantizer import _native as N
    C, K, g = 24, 2048, 128
    w = datagen.weights((C, K), "bf16", 99).float()
    w[0, :128] = 0.0
    w[1, 128:256] = 0.5
    w[2, 5] = float("nan")
    w[3, 300] = float("inf")
    w[4, 700] = float("-inf")
    w[5] *= 1e30
    w[6] *= 1e-30
    w[7] = w[7].abs() + 0.25
    w[8] = -w[8].abs() - 0.25
    w[9, 1024:1152] = torch.linspace(-1, 1, 128)
    w[10, :] = 3.0e38
    w = w.to(torch.bfloat16)
    s = _scales(K, 5)
    s[::7] = 2.0 ** -20
    s[3::11] = 2.0 ** 20
    want = O.pack_result(O.quantize_scaled(w, s, 4, g, sym))
    _check(_run(native_lib, N, w, s, g, sym, cuda_device), want, "special")


def test_colscale_many_row_blocks_and_canaries(native_lib, cuda_device):
    """more row blocks than CTAs per slab (several pipeline iterations per CTA), ragged last row block,
    guarded output buffers"""
    from awq_quantizer import _native as N
    from tests.test_gpu_parity import _canaries_intact, _guarded
    dev = cuda_device
    C, K, g = 8 * 700 + 3, 1024, 128
    G = K // g
    w = datagen.weights((C, K), "bf16", 17)
    s = _scales(K, 3)
    want = O.pack_result(O.quantize_scaled(w, s, 4, g, False))
    sizes = {"q": C * K * 4, "qp": C * K // 2, "s": C * G * 2, "z": C * G * 4, "zq": C * (G // 8) * 4}
    bufs = {k: _guarded(v, dev) for k, v in sizes.items()}
    wd, sd = w.to(dev), s.to(dev)
    rc = native_lib.awqk_group_quant(wd.data_ptr(), N.BF16, C, K, g, 4, 0, N.ARITH_FP32, bufs["q"][1].data_ptr(),
                                     bufs["qp"][1].data_ptr(), bufs["s"][1].data_ptr(), bufs["z"][1].data_ptr(),
                                     bufs["zq"][1].data_ptr(), sd.data_ptr(), None)
    assert rc == 0
    torch.cuda.synchronize()
    for k, (b, _) in bufs.items():
        assert _canaries_intact(b, sizes[k]), k
    assert_same(bufs["q"][1].view(torch.int32).reshape(C, K).cpu(), want["tensor_q"], "q")
    assert_same(bufs["qp"][1].view(torch.int32).reshape(C, K // 8).cpu(), want["qweight"], "qweight")
    assert_same(bufs["s"][1].view(torch.float16).reshape(C, G).cpu(), want["scales"], "scales")
    assert_same(bufs["z"][1].view(torch.int32).reshape(C, G).cpu(), want["zero_points"], "zp")
    assert_same(bufs["zq"][1].view(torch.int32).reshape(C, G // 8).cpu(), want["qzeros"], "qzeros")


def test_colscale_matches_register_path(native_lib, cuda_device, monkeypatch):
    """the slab kernel and the register-path kernel (K not a multiple of 1024 -> group_quant_flat) agree
    on the shared left part of a matrix whose column scales are equal there"""
    from awq_quantizer import _native as N
    C, g = 64, 128
    w = datagen.weights((C, 2048), "bf16", 123)
    s = _scales(2048, 9)
    a = _run(native_lib, N, w, s, g, False, cuda_device)                                   # slab kernel
    b = _run(native_lib, N, w[:, :1536].contiguous(), s[:1536].contiguous(), g, False, cuda_device)   # register path
    assert torch.equal(a["tensor_q"][:, :1536], b["tensor_q"])
    assert torch.equal(a["scales"][:, :12], b["scales"])
    assert torch.equal(a["zero_points"][:, :12], b["zero_points"])


def _batch_case(native_lib, N, dev, shapes, dt, g, sym, *, unpacked, guard=False):
    """awqk_group_quant_batch over `shapes` vs the oracle, tensor by tensor"""
    from tests.test_gpu_parity import _canaries_intact, _guarded
    items, keep, wants = [], [], []
    for i, (C, K) in enumerate(shapes):
        G = K // g
        w = datagen.weights((C, K), dt, datagen.seed_of("csb", i, C, K, dt))
        s = _scales(K, 31 * i + K)
        wants.append(O.pack_result(O.quantize_scaled(w, s, 4, g, sym)))
        sizes = {"q": C * K * 4, "qp": C * K // 2, "s": C * G * 2, "z": C * G * 4, "zq": C * -(-G // 8) * 4}
        bufs = {k: _guarded(v, dev) for k, v in sizes.items()}
        wd, sd = w.to(dev), s.to(dev)
        keep.append((wd, sd, bufs, sizes, (C, K, G)))
        items.append((wd, C, K, sd, bufs["q"][1] if unpacked else None, bufs["qp"][1], bufs["s"][1], bufs["z"][1], bufs["zq"][1]))
    N.group_quant_batch(items, N.dtype_code(keep[0][0].dtype), g, 4, sym, N.ARITH_FP32, None)
    torch.cuda.synchronize()
    for i, ((wd, sd, bufs, sizes, (C, K, G)), want) in enumerate(zip(keep, wants)):
        what = f"item {i} {C}x{K}"
        for k, (b, _) in bufs.items():
            if k == "q" and not unpacked:
                continue
            assert _canaries_intact(b, sizes[k]), f"{what}/{k} canary"
        if unpacked:
            assert_same(bufs["q"][1].view(torch.int32).reshape(C, K).cpu(), want["tensor_q"], what + "/q")
        assert_same(bufs["qp"][1].view(torch.int32).reshape(C, K // 8).cpu(), want["qweight"], what + "/qweight")
        assert_same(bufs["s"][1].view(torch.float16).reshape(C, G).cpu(), want["scales"], what + "/scales")
        assert_same(bufs["z"][1].view(torch.int32).reshape(C, G).cpu(), want["zero_points"], what + "/zp")
        if G % 8 == 0:
            assert_same(bufs["zq"][1].view(torch.int32).reshape(C, G // 8).cpu(), want["qzeros"], what + "/qzeros")


@pytest.mark.parametrize("unpacked", [False, True])
@pytest.mark.parametrize("dt,g,sym", [("bf16", 128, False), ("fp16", 64, True), ("bf16", 32, False)])
def test_colscale_batch_bit_exact(native_lib, cuda_device, dt, g, sym, unpacked):
    """one launch over tensors of different heights and widths: units of several tensors interleave on a CTA,
    ragged last row chunks (C % 8 != 0, C < 8), single-slab and many-slab tensors, guarded outputs"""
    from awq_quantizer import _native as N
    shapes = [(300, 4096), (1, 1024), (77, 3072), (8, 1024), (1029, 2048), (13, 7168), (5, 1024)]
    _batch_case(native_lib, N, cuda_device, shapes, dt, g, sym, unpacked=unpacked)


def test_colscale_batch_many_units_per_cta(native_lib, cuda_device):
    """enough rows that every persistent CTA walks through many units (table double buffering, both parities of
    the table barriers, tensors changing in the middle of a CTA's sequence)"""
    from awq_quantizer import _native as N
    shapes = [(8 * 700 + 3, 1024), (4099, 2048), (2500, 1024), (3001, 3072)]
    _batch_case(native_lib, N, cuda_device, shapes, "bf16", 128, False, unpacked=False)


def test_colscale_batch_more_than_one_launch_and_fallbacks(native_lib, cuda_device):
    """> 32 eligible tensors (two launches) mixed with tensors the slab kernel does not take (K % 1024 != 0:
    register path inside the same call)"""
    from awq_quantizer import _native as N
    shapes = [(16 + i, 1024 * (1 + i % 3)) for i in range(37)] + [(40, 1536), (9, 512)]
    _batch_case(native_lib, N, cuda_device, shapes, "bf16", 128, False, unpacked=False)


@pytest.mark.parametrize("want_zp,want_zq", [(False, True), (True, False), (False, False)])
def test_colscale_batch_output_subsets(native_lib, cuda_device, want_zp, want_zq):
    """launch-uniform output flags: packed zeros only (what the packed model path asks for), int32 zero points only,
    neither -- and a batch that mixes tensors with different output sets (split into one launch per set)"""
    from awq_quantizer import _native as N
    dev = cuda_device
    g = 128
    shapes = [(40, 2048), (9, 1024), (130, 3072)]
    items, keep, wants = [], [], []
    for i, (C, K) in enumerate(shapes):
        G = K // g
        w = datagen.weights((C, K), "bf16", datagen.seed_of("csf", i, C, K))
        s = _scales(K, 7 * i + 1)
        wants.append(O.pack_result(O.quantize_scaled(w, s, 4, g, False)))
        mixed_zp = want_zp or (i == 2 and not want_zq)            # third tensor of the (False, False) case differs on purpose
        o = {"qp": torch.full((C, K // 8), 77, dtype=torch.int32, device=dev),
             "s": torch.zeros((C, G), dtype=torch.float16, device=dev),
             "z": torch.full((C, G), 77, dtype=torch.int32, device=dev) if mixed_zp else None,
             "zq": torch.full((C, G // 8), 77, dtype=torch.int32, device=dev) if want_zq else None}
        keep.append((w.to(dev), s.to(dev), o))
        items.append((keep[-1][0], C, K, keep[-1][1], None, o["qp"], o["s"], o["z"], o["zq"]))
    N.group_quant_batch(items, N.BF16, g, 4, False, N.ARITH_FP32, None)
    torch.cuda.synchronize()
    for (w, s, o), want in zip(keep, wants):
        assert_same(o["qp"].cpu(), want["qweight"], "qweight")
        assert_same(o["s"].cpu(), want["scales"], "scales")
        if o["z"] is not None:
            assert_same(o["z"].cpu(), want["zero_points"], "zero_points")
        if o["zq"] is not None:
            assert_same(o["zq"].cpu(), want["qzeros"], "qzeros")


def test_colscale_batch_equals_single_calls(native_lib, cuda_device):
    """Llama-3-8B layer shapes (q, k, v, o, gate, up, down) in one launch == seven awqk_group_quant calls"""
    from awq_quantizer import _native as N
    dev = cuda_device
    shapes = [(4096, 4096), (1024, 4096), (1024, 4096), (4096, 4096), (14336, 4096), (14336, 4096), (4096, 14336)]
    g = 128
    single, items, outs = [], [], []
    for i, (C, K) in enumerate(shapes):
        G = K // g
        gen = torch.Generator(device=dev).manual_seed(1000 + i)
        w = (torch.randn((C, K), generator=gen, device=dev) * 0.02).to(torch.bfloat16)
        s = torch.exp(0.5 * torch.randn(K, generator=gen, device=dev)).float()
        mk = lambda: {"qp": torch.full((C, K // 8), 77, dtype=torch.int32, device=dev), "s": torch.zeros((C, G), dtype=torch.float16, device=dev),
                      "z": torch.full((C, G), 77, dtype=torch.int32, device=dev), "zq": torch.full((C, G // 8), 77, dtype=torch.int32, device=dev)}
        a, b = mk(), mk()
        N.check(native_lib.awqk_group_quant(w.data_ptr(), N.BF16, C, K, g, 4, 0, N.ARITH_FP32, None, a["qp"].data_ptr(),
                                            a["s"].data_ptr(), a["z"].data_ptr(), a["zq"].data_ptr(), s.data_ptr(), None))
        single.append(a)
        outs.append(b)
        items.append((w, C, K, s, None, b["qp"], b["s"], b["z"], b["zq"]))
    N.group_quant_batch(items, N.BF16, g, 4, False, N.ARITH_FP32, None)
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(zip(single, outs)):
        for k in a:
            assert torch.equal(a[k].view(torch.int32) if k != "s" else a[k].view(torch.int16),
                               b[k].view(torch.int32) if k != "s" else b[k].view(torch.int16)), (i, k)


def test_colscale_batch_randomized_vs_register_path(native_lib, cuda_device):
    """randomized batches (1..31 tensors, heights 1..3000, 1..8 slabs, every group size, both symmetries) against the
    independent register-path kernel (group_quant_flat with col_scale), which awqk_group_quant selects when q_packed is
    only 8-byte aligned -- two implementations of the same arithmetic, compared on ~10^8 elements"""
    import random
    from awq_quantizer import _native as N
    dev = cuda_device
    rnd = random.Random(2024)
    for case in range(12):
        g = rnd.choice([32, 64, 128])
        sym = rnd.random() < 0.3
        n = rnd.choice([1, 2, 5, 17, 31])
        items, ref = [], []
        for i in range(n):
            C = rnd.choice([1, 7, 8, 9, 63, 250, rnd.randint(1, 3000)])
            K = 1024 * rnd.randint(1, 8)
            G = K // g
            gen = torch.Generator(device=dev).manual_seed(1000 * case + i)
            w = (torch.randn((C, K), generator=gen, device=dev) * (0.02 if i % 3 else 3.0)).to(torch.bfloat16)
            if i % 5 == 0:
                w[:, :g] = w[:, :g].abs() + 0.01                       # all-positive groups: clamped zero points
            s = torch.exp(0.7 * torch.randn(K, generator=gen, device=dev)).float()
            out = {"qp": torch.zeros((C, K // 8), dtype=torch.int32, device=dev), "s": torch.zeros((C, G), dtype=torch.float16, device=dev),
                   "z": torch.zeros((C, G), dtype=torch.int32, device=dev), "zq": torch.zeros((C, -(-G // 8)), dtype=torch.int32, device=dev)}
            items.append((w, C, K, s, None, out["qp"], out["s"], out["z"], out["zq"]))
            # reference: same call with a q_packed pointer that is 8- but not 16-byte aligned -> register path
            buf = torch.zeros(C * (K // 8) + 4, dtype=torch.int32, device=dev)
            qp = buf[2:2 + C * (K // 8)].view(C, K // 8)
            assert qp.data_ptr() % 16 == 8
            r = {"qp": qp, "s": torch.zeros_like(out["s"]), "z": torch.zeros_like(out["z"]), "zq": torch.zeros_like(out["zq"])}
            N.check(native_lib.awqk_group_quant(w.data_ptr(), N.BF16, C, K, g, 4, int(sym), N.ARITH_FP32, None, qp.data_ptr(),
                                                r["s"].data_ptr(), r["z"].data_ptr(), r["zq"].data_ptr(), s.data_ptr(), None))
            ref.append((out, r))
        N.group_quant_batch(items, N.BF16, g, 4, sym, N.ARITH_FP32, None)
        torch.cuda.synchronize()
        for i, (out, r) in enumerate(ref):
            for k in ("qp", "z", "zq"):
                assert torch.equal(out[k], r[k]), (case, i, k, items[i][1:3], g, sym)
            assert torch.equal(out["s"].view(torch.int16), r["s"].view(torch.int16)), (case, i, "scales")

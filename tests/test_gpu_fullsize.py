"""GPU (B200): parity at BASELINE.json's FULL tensor sizes, driver-run (these were builder-run tools in round 1:
tools/verify_search_parity.py, tools/verify_model_parity.py).

* the activation-aware search on real Llama-3-8B linear shapes, T = 2048, 20-point grid, against the oracle's
  definition (PARITY UNPINNED for the search itself: the reference has none; the oracle composes the reference's
  pinned group quantizer).  Bars (north star): alpha exact given equal scales, error scores within 1e-3 relative,
  final qweight / qzeros / scales / tensor_q bit-exact.
* K1 against the C restatement of the reference (oracle/awq_oracle.c, pinned through tests/golden) on more than
  10^9 elements of the Llama-3-8B shape, every output bit-for-bit.
* one C-ABI call awqk_scale_search == the stand-alone stages (delta -> HBM -> GEMM) + host argmin + K1.
"""
import ctypes
import zlib

import pytest
import torch

from oracle import awq_oracle as O
from tests import datagen
from tests.util import assert_same

pytestmark = pytest.mark.gpu


def _bits(t):
    return t.view(torch.int16) if t.dtype == torch.float16 else t


@pytest.mark.timeout(900)
@pytest.mark.parametrize("C,K", [(4096, 4096), (14336, 4096), (4096, 14336)])
def test_search_full_size_vs_oracle(native_lib, cuda_device, C, K):
    from awq_quantizer.quantization import AWQQuantizer
    from awq_quantizer.quantization.search import search_device
    T, n = 2048, 20
    W = datagen.weights((C, K), "bf16", datagen.seed_of("vs", C, K))
    X = datagen.activations(T, K, "bf16", datagen.seed_of("vsx", K))
    qz = AWQQuantizer(bits=4, group_size=128, symmetric=False, device="cuda:0", logger_level="ERROR", n_grid=n)
    got = qz.quantize(W, activations=X, pack=True)
    r = search_device(W.to(cuda_device), X.to(cuda_device), bits=4, group_size=128, symmetric=False, n_grid=n)
    s_grid = r["s_grid"].cpu()
    want = O.search_scales(W, X, 4, 128, False, n_grid=n, s_grid=s_grid)     # "given equal scales"
    rel = max(abs(float(got["search_err"][i]) - want["err"][i]) / want["err"][i] for i in range(n))
    assert rel <= 1e-3, rel
    srt = sorted(want["err"])
    assert (srt[1] - srt[0]) > 1e-3 * srt[0], "test data must not sit on a near-tie"    # so that alpha IS asserted
    assert int(got["best_idx"]) == want["best_idx"] and abs(float(got["alpha"]) - want["alpha"]) < 1e-7
    assert int(got["best_idx"]) > 0
    assert torch.equal(got["awq_scale"], s_grid[int(got["best_idx"])])
    final = O.pack_result(O.quantize_scaled(W, got["awq_scale"], 4, 128, False))
    for k in ("tensor_q", "scales", "zero_points", "qweight", "qzeros"):
        assert torch.equal(_bits(got[k]), _bits(final[k])), k


@pytest.mark.timeout(900)
def test_k1_over_1e9_elements_vs_c_oracle(native_lib, cuda_device):
    """embed_tokens + the seven linears of five layers of the Llama-3-8B shape: 1.6e9 elements, every output of K1
    (int32 codes, fp16 scales, int32 zero points, packed words) against the C restatement of the reference"""
    from awq_quantizer import model_shapes as M
    from awq_quantizer.quantization import AWQQuantizer
    from oracle import c_oracle as CO
    qz = AWQQuantizer(bits=4, group_size=128, symmetric=False, device="cuda:0", logger_level="ERROR")
    gen = torch.Generator(device=cuda_device)
    specs = [(n, s) for n, s, _ in M.workload("llama3-8b")
             if n == "model.embed_tokens.weight" or any(f"layers.{i}." in n for i in range(5))]
    done = 0
    for name, shape in specs:
        gen.manual_seed(zlib.crc32(name.encode()) ^ 0xA11CE)
        w = (torch.randn(shape, generator=gen, device=cuda_device, dtype=torch.float32) * 0.02).to(torch.bfloat16)
        r = {k: v.cpu() for k, v in qz._quantize_device(w, pack=True, unpacked=True).items()}
        wc = w.cpu()
        want = CO.group_quant(wc, 4, 128, False)
        Cn = shape[0] if len(shape) > 1 else 1
        assert torch.equal(r["tensor_q"], want["tensor_q"]), name
        assert torch.equal(r["zero_points"], want["zero_points"].reshape(r["zero_points"].shape)), name
        assert torch.equal(_bits(r["scales"]), _bits(want["scales"]).reshape(r["scales"].shape)), name
        assert torch.equal(r["qweight"], CO.pack_rows(want["tensor_q"].reshape(Cn, -1), 0, 4)), name
        assert torch.equal(r["qzeros"], CO.pack_rows(want["zero_points"].reshape(Cn, -1), 0, 4)), name
        done += wc.numel()
        del w, r, want
    assert done > 1_000_000_000, done


@pytest.mark.parametrize("dt,g,sym,C,K,T", [("bf16", 128, False, 520, 1024, 300), ("fp16", 64, True, 256, 2048, 257),
                                             ("fp32", 32, False, 300, 512, 128), ("bf16", 128, True, 1024, 4096, 1000),
                                             # ragged last panel (K = 2048 + 192), fewer rows than one slab
                                             ("bf16", 64, False, 40, 2240, 200),
                                             # more m-tiles than CTA pairs (a slab spans several waves), 3 panels
                                             ("bf16", 128, False, 300, 4224, 19500)])
def test_scale_search_one_call_c_abi(native_lib, cuda_device, dt, g, sym, C, K, T):
    """awqk_scale_search called directly (own grid from X inside the workspace, minimum workspace): scores ==
    stand-alone delta + GEMM stages, best = first minimum, best_s = that grid row, outputs == awqk_group_quant"""
    from awq_quantizer import _native as N
    L = native_lib
    dev = cuda_device
    n = 12
    W = datagen.weights((C, K), dt, datagen.seed_of("ss", dt, g, C)).to(dev)
    X = datagen.activations(T, K, "bf16", datagen.seed_of("ssx", K)).to(dev)
    mn = ctypes.c_size_t(0)
    pref = L.awqk_workspace_bytes(C, K, T, n, 0, ctypes.byref(mn))
    assert 0 < mn.value <= pref
    assert L.awqk_workspace_bytes(0, K, T, n, 0, None) == 0
    G = K // g
    for nbytes in (mn.value, pref):
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        err = torch.empty(n, dtype=torch.float64, device=dev)
        best = torch.empty((), dtype=torch.int32, device=dev)
        s_best = torch.empty(K, dtype=torch.float32, device=dev)
        qw = torch.empty((C, K // 8), dtype=torch.int32, device=dev)
        q = torch.empty((C, K), dtype=torch.int32, device=dev)
        sc = torch.empty((C, G), dtype=torch.float16, device=dev)
        zp = torch.empty((C, G), dtype=torch.int32, device=dev)
        zq = torch.empty((C, -(-G // 8)), dtype=torch.int32, device=dev)
        rc = L.awqk_scale_search(W.data_ptr(), N.dtype_code(W.dtype), C, K, X.data_ptr(), T, None, n, g, 4, int(sym),
                                 err.data_ptr(), best.data_ptr(), s_best.data_ptr(), q.data_ptr(), qw.data_ptr(),
                                 sc.data_ptr(), zp.data_ptr(), zq.data_ptr(), ws.data_ptr(), nbytes, None)
        assert rc == 0, rc
        torch.cuda.synchronize()
        # the same grid through the stand-alone calls
        colsum = torch.zeros(K, dtype=torch.float64, device=dev)
        grid = torch.empty((n, K), dtype=torch.float32, device=dev)
        w2 = torch.empty(2 * n, dtype=torch.float32, device=dev)
        assert L.awqk_abs_colsum(X.data_ptr(), N.BF16, T, K, colsum.data_ptr(), None) == 0
        assert L.awqk_alpha_grid(colsum.data_ptr(), T, K, n, grid.data_ptr(), w2.data_ptr(), None) == 0
        dw = torch.empty((n, C, K), dtype=torch.bfloat16, device=dev)
        e2 = torch.zeros(n, dtype=torch.float64, device=dev)
        assert L.awqk_fakequant_delta(W.data_ptr(), N.dtype_code(W.dtype), C, K, g, 4, int(sym), grid.data_ptr(), n,
                                      dw.data_ptr(), None) == 0
        assert L.awqk_sqerr_gemm(X.data_ptr(), dw.data_ptr(), T, C, K, n, e2.data_ptr(), None) == 0
        torch.cuda.synchronize()
        want = (e2 / float(T * C)).cpu()
        assert torch.allclose(err.cpu(), want, rtol=1e-9, atol=0)          # same products; fp64 atomic order only
        b = int(torch.argmin(want))
        assert int(best) == b
        assert torch.equal(s_best.cpu(), grid[b].cpu())
        q2, qw2, sc2, zp2, zq2 = (torch.empty_like(t) for t in (q, qw, sc, zp, zq))
        assert L.awqk_group_quant(W.data_ptr(), N.dtype_code(W.dtype), C, K, g, 4, int(sym), N.ARITH_FP32, q2.data_ptr(),
                                  qw2.data_ptr(), sc2.data_ptr(), zp2.data_ptr(), zq2.data_ptr(), s_best.data_ptr(), None) == 0
        torch.cuda.synchronize()
        for a, c, what in ((q, q2, "q"), (qw, qw2, "qweight"), (sc, sc2, "scales"), (zp, zp2, "zp"), (zq, zq2, "qzeros")):
            assert_same(a.cpu(), c.cpu(), what)
    # argument errors: too small a workspace, misaligned workspace, bad group size
    small = torch.empty(mn.value, dtype=torch.uint8, device=dev)
    args = [W.data_ptr(), N.dtype_code(W.dtype), C, K, X.data_ptr(), T, None, n, g, 4, int(sym), None, None,
            s_best.data_ptr(), None, None, None, None, None]
    assert L.awqk_scale_search(*args, small.data_ptr(), mn.value - 256, None) == -5
    assert L.awqk_scale_search(*args, small.data_ptr() + 8, mn.value - 8, None) == -2
    args[8] = 48
    assert L.awqk_scale_search(*args, small.data_ptr(), mn.value, None) == -4

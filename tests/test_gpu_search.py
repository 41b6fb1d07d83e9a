"""GPU (B200): the activation-aware search kernels (K2) against the oracle's definition
(oracle/awq_oracle.py::search_scales -- PARITY UNPINNED: the reference has no such code, the oracle
composes the reference's pinned group quantizer).

Tolerances (north star): chosen alpha exact and final qweight / qzeros bit-exact *given equal scales*
(the GPU's s grid is injected into the oracle), error scores within 1e-3 relative."""
import pytest
import torch

from oracle import awq_oracle as O
from tests import datagen
from tests.util import assert_quant_equal, assert_same

pytestmark = pytest.mark.gpu


def test_abs_colsum_exact(native_lib, cuda_device):
    from awq_quantizer import _native as N
    for dt in ("bf16", "fp16", "fp32"):
        X = datagen.activations(300, 520, dt, 11)
        xd = X.to(cuda_device)
        cs = torch.zeros(520, dtype=torch.float64, device=cuda_device)
        assert native_lib.awqk_abs_colsum(xd.data_ptr(), N.dtype_code(xd.dtype), 300, 520, cs.data_ptr(), None) == 0
        torch.cuda.synchronize()
        want = X.double().abs().sum(0)
        assert torch.equal(cs.cpu(), want), dt                     # fp64 accumulation of fp16/bf16 data is exact
        assert torch.equal((cs.cpu() / 300).float(), O.activation_mean(X))


def test_alpha_grid_close_to_oracle(native_lib, cuda_device):
    from awq_quantizer import _native as N
    X = datagen.activations(256, 1024, "bf16", 5)
    m = O.activation_mean(X)
    cs = X.double().abs().sum(0).to(cuda_device)
    n = 20
    s = torch.empty((n, 1024), dtype=torch.float32, device=cuda_device)
    ws = torch.empty(2 * n, dtype=torch.float32, device=cuda_device)
    assert native_lib.awqk_alpha_grid(cs.data_ptr(), 256, 1024, n, s.data_ptr(), ws.data_ptr(), None) == 0
    torch.cuda.synchronize()
    s = s.cpu()
    assert torch.equal(s[0], torch.ones(1024))                     # alpha = 0
    for i in range(n):
        want = O.alpha_scales(m, i / n)
        assert torch.allclose(s[i], want, rtol=2e-6, atol=0), i    # powf ulps; everything else is exact
        assert abs(float(s[i].max() * s[i].min()) - 1.0) < 1e-5    # normalisation: max * min == 1


@pytest.mark.parametrize("dt", ["bf16", "fp16", "fp32"])
@pytest.mark.parametrize("sym", [False, True])
@pytest.mark.parametrize("g", [32, 128])
def test_fakequant_delta_bit_exact(native_lib, cuda_device, dt, sym, g):
    from awq_quantizer import _native as N
    C, K, n_s = 48, 512, 3
    W = datagen.weights((C, K), dt, datagen.seed_of("fq", dt, sym, g))
    W[1, :128] = 0.0
    X = datagen.activations(64, K, "bf16", 8)
    m = O.activation_mean(X)
    s = torch.stack([O.alpha_scales(m, a) for a in (0.0, 0.35, 0.8)])
    wd, sd = W.to(cuda_device), s.to(cuda_device)
    dw = torch.zeros((n_s, C, K), dtype=torch.bfloat16, device=cuda_device)
    assert native_lib.awqk_fakequant_delta(wd.data_ptr(), N.dtype_code(wd.dtype), C, K, g, 4, int(sym),
                                           sd.data_ptr(), n_s, dw.data_ptr(), None) == 0
    torch.cuda.synchronize()
    for i in range(n_s):
        want = O.fake_quant_delta(W, s[i], 4, g, sym).to(torch.bfloat16)
        assert_same(dw[i].cpu().view(torch.int16), want.view(torch.int16), f"dW[{i}]")


@pytest.mark.parametrize("T,C,K,n_s", [(128, 256, 64, 1), (256, 512, 256, 3), (200, 300, 320, 2), (512, 1024, 1024, 5),
                                       (130, 260, 72, 1)])
def test_sqerr_gemm_vs_fp64(native_lib, cuda_device, T, C, K, n_s):
    """the tcgen05 GEMM + fused sum-of-squares epilogue on arbitrary bf16 operands (incl. tiles that run
    past T, C and K: TMA zero fill)"""
    g = torch.Generator().manual_seed(T * 7 + C)
    X = (torch.randn((T, K), generator=g)).to(torch.bfloat16)
    D = (torch.randn((n_s, C, K), generator=g) * 0.01).to(torch.bfloat16)
    xd, dd = X.to(cuda_device), D.to(cuda_device)
    err = torch.zeros(n_s, dtype=torch.float64, device=cuda_device)
    assert native_lib.awqk_sqerr_gemm(xd.data_ptr(), dd.data_ptr(), T, C, K, n_s, err.data_ptr(), None) == 0
    torch.cuda.synchronize()
    for i in range(n_s):
        want = float(((X.double() @ D[i].double().T) ** 2).sum())
        got = float(err[i])
        assert abs(got - want) <= 1e-5 * want, (i, got, want)
    # accumulating call: err += ...
    assert native_lib.awqk_sqerr_gemm(xd.data_ptr(), dd.data_ptr(), T, C, K, n_s, err.data_ptr(), None) == 0
    torch.cuda.synchronize()
    want0 = float(((X.double() @ D[0].double().T) ** 2).sum())
    assert abs(float(err[0]) - 2 * want0) <= 2e-5 * want0


def test_sqerr_gemm_persistent_many_tiles(native_lib, cuda_device):
    """more tiles than SMs and many k-blocks: exercises ring wrap-around and both TMEM buffers"""
    T, C, K, n_s = 1024, 2048, 2048, 6
    g = torch.Generator(device=cuda_device).manual_seed(3)
    X = torch.randn((T, K), generator=g, device=cuda_device).to(torch.bfloat16)
    D = (torch.randn((n_s, C, K), generator=g, device=cuda_device) * 0.01).to(torch.bfloat16)
    err = torch.zeros(n_s, dtype=torch.float64, device=cuda_device)
    assert native_lib.awqk_sqerr_gemm(X.data_ptr(), D.data_ptr(), T, C, K, n_s, err.data_ptr(), None) == 0
    torch.cuda.synchronize()
    for i in range(n_s):
        want = float(((X.double() @ D[i].double().T) ** 2).sum())   # fp64 matmul on the GPU (checker only)
        assert abs(float(err[i]) - want) <= 1e-5 * want, (i, float(err[i]), want)


@pytest.mark.parametrize("sym", [False, True])
def test_full_search_vs_oracle(native_lib, cuda_device, sym):
    from awq_quantizer.quantization import AWQQuantizer
    C, K, T, n = 256, 512, 192, 20
    W = datagen.weights((C, K), "bf16", 21)
    X = datagen.activations(T, K, "bf16", 22)
    qz = AWQQuantizer(bits=4, group_size=128, symmetric=sym, device="cuda:0", logger_level="ERROR", n_grid=n)
    got = qz.quantize(W, activations=X, pack=True)
    # the GPU's scale grid is recomputed here and injected into the oracle ("given equal scales")
    from awq_quantizer.quantization.search import search_device
    r = search_device(W.to(cuda_device), X.to(cuda_device), bits=4, group_size=128, symmetric=sym, n_grid=n)
    s_grid = r["s_grid"].cpu()
    want = O.search_scales(W, X, 4, 128, sym, n_grid=n, s_grid=s_grid)
    err = got["search_err"]
    for i in range(n):
        assert abs(float(err[i]) - want["err"][i]) <= 1e-3 * want["err"][i], (i, float(err[i]), want["err"][i])
    srt = sorted(want["err"])
    near_tie = (srt[1] - srt[0]) < 1e-3 * srt[0]
    if not near_tie:                                              # near-ties are reported, not failed
        assert int(got["best_idx"]) == want["best_idx"]
        assert abs(float(got["alpha"]) - want["alpha"]) < 1e-7
    assert int(got["best_idx"]) > 0                               # activation-aware scaling beats alpha = 0 here
    s_best = got["awq_scale"]
    assert torch.equal(s_best, s_grid[int(got["best_idx"])])
    final = O.pack_result(O.quantize_scaled(W, s_best, 4, 128, sym))
    assert_quant_equal(got, final, "final", keys=("tensor_q", "scales", "zero_points", "qweight", "qzeros"))
    # scales from the oracle's own s (CPU pow) stay within 1 fp16 ulp
    own = O.quantize_scaled(W, want["s_grid"][int(got["best_idx"])], 4, 128, sym)
    a, b = got["scales"].view(torch.int16).int(), own["scales"].view(torch.int16).int()
    assert int((a - b).abs().max()) <= 1


def test_search_argument_errors(native_lib, cuda_device):
    from awq_quantizer.quantization import AWQQuantizer
    qz = AWQQuantizer(bits=4, group_size=128, symmetric=False, device="cuda:0", logger_level="ERROR")
    W = datagen.weights((64, 256), "bf16", 1)
    with pytest.raises(ValueError):
        qz.quantize(W, activations=torch.zeros(16, 128, dtype=torch.bfloat16))     # wrong K
    with pytest.raises(ValueError):
        qz.quantize(W.reshape(64, 2, 128), activations=torch.zeros(16, 256, dtype=torch.bfloat16))  # not 2-D
    assert native_lib.awqk_sqerr_gemm(None, None, 1, 1, 8, 1, None, None) == -1


def test_search_pipeline_matches_per_tensor_search(native_lib, cuda_device):
    """SearchPipeline (two streams, batched argmin) == search_device per tensor"""
    from awq_quantizer.quantization.search import SearchPipeline, search_device
    dev = cuda_device
    X1 = datagen.activations(160, 256, "bf16", 1).to(dev)
    X2 = datagen.activations(96, 512, "bf16", 2).to(dev)
    ws = {"a": (datagen.weights((64, 256), "bf16", 3).to(dev), X1), "b": (datagen.weights((128, 256), "bf16", 4).to(dev), X1),
          "c": (datagen.weights((32, 512), "bf16", 5).to(dev), X2), "d": (datagen.weights((256, 256), "fp16", 6).to(dev), X1)}
    pipe = SearchPipeline(dev, bits=4, group_size=128, symmetric=False, n_grid=8)
    for rep in range(2):                      # second round reuses events / buffers
        for n, (w, x) in ws.items():
            pipe.submit(n, w, x)
        res = {name: (mean, best, sb) for name, mean, best, sb in pipe.finish()}
        torch.cuda.synchronize()
        assert list(res) == list(ws)
        for n, (w, x) in ws.items():
            r = search_device(w, x, bits=4, group_size=128, symmetric=False, n_grid=8)
            want = (r["err_sum"] / float(x.shape[0] * w.shape[0])).cpu()
            got = res[n][0].cpu()
            assert torch.allclose(got, want, rtol=1e-9, atol=0), n       # same kernels; fp64 atomics order only
            b = int(torch.argmin(want))
            assert int(res[n][1]) == b
            assert torch.equal(res[n][2].cpu(), r["s_grid"][b].cpu())


def test_quantize_model_with_activations(native_lib, cuda_device):
    """model-level AWQ through the public API: searched linears + plain tensors in one call"""
    from awq_quantizer.quantization import AWQQuantizer
    X1 = datagen.activations(128, 256, "bf16", 1)
    X2 = datagen.activations(96, 512, "bf16", 2)
    tensors = {"q_proj": datagen.weights((128, 256), "bf16", 3), "k_proj": datagen.weights((64, 256), "bf16", 4),
               "norm": datagen.weights((256,), "bf16", 5), "down": datagen.weights((64, 512), "bf16", 6)}
    acts = {"q_proj": X1, "k_proj": X1, "down": X2}
    qz = AWQQuantizer(bits=4, group_size=128, symmetric=False, device="cuda:0", logger_level="ERROR", n_grid=10)
    out = qz.quantize_model(tensors, activations=acts, pack=True, )
    both = qz.quantize_model(tensors, activations=acts, pack=True, keep_unpacked=True)
    plain = qz.quantize_model(tensors, activations=acts)
    assert list(out) == list(tensors)
    for n in acts:
        single = qz.quantize(tensors[n], activations=acts[n], pack=True)
        assert int(out[n]["best_idx"]) == int(single["best_idx"])
        assert abs(float(out[n]["alpha"]) - float(single["alpha"])) == 0
        assert torch.equal(out[n]["awq_scale"], single["awq_scale"])
        assert "tensor_q" not in out[n] and "qweight" not in plain[n]      # packed results skip the 4 B/element codes
        assert_quant_equal(out[n], single, n, keys=("scales", "zero_points", "qweight", "qzeros"))
        assert_quant_equal(both[n], single, n, keys=("tensor_q", "scales", "zero_points", "qweight", "qzeros"))
        assert_quant_equal(plain[n], single, n, keys=("tensor_q", "scales", "zero_points"))
        assert torch.allclose(out[n]["search_err"], single["search_err"], rtol=1e-9)
    want = O.pack_result(O.group_quant_vec(tensors["norm"], 4, 128, False, True))
    assert_same(out["norm"]["qweight"], want["qweight"], "norm")
    assert "awq_scale" not in out["norm"]


def test_streamed_model_search_many_waves(native_lib, cuda_device):
    """quantize_model_with_search with waves far smaller than the model: slot reuse in the uploader (pinned
    staging + device slots), pageable / pinned / device-resident inputs mixed, results identical to the
    one-tensor API"""
    from awq_quantizer.quantization import AWQQuantizer
    from awq_quantizer.quantization.search import quantize_model_with_search
    X = {256: datagen.activations(96, 256, "bf16", 1), 512: datagen.activations(64, 512, "bf16", 2)}
    shapes = [(64, 256), (128, 512), (32, 256), (96, 512), (64, 512), (16, 256), (48, 256), (8, 512), (72, 256)]
    tensors, acts = {}, {}
    for i, s in enumerate(shapes):
        w = datagen.weights(s, "fp16" if i == 4 else "bf16", 100 + i)
        if i % 3 == 1:
            w = w.pin_memory()
        elif i % 3 == 2:
            w = w.to(cuda_device)
        tensors[f"t{i}"] = w
        acts[f"t{i}"] = X[s[1]]
    qz = AWQQuantizer(bits=4, group_size=128, symmetric=False, device="cuda:0", logger_level="ERROR", n_grid=6)
    for wave_bytes, pin in ((1, False), (70_000, False), (70_000, True), (1 << 30, False)):
        out = quantize_model_with_search(qz, tensors, acts, cuda_device, pack=True, wave_bytes=wave_bytes, pin_results=pin)
        assert list(out) == list(tensors)
        for n, w in tensors.items():
            single = qz.quantize(w.cpu(), activations=acts[n], pack=True)
            assert int(out[n]["best_idx"]) == int(single["best_idx"]), (n, wave_bytes)
            assert_quant_equal(out[n], single, f"{n}/{wave_bytes}", keys=("scales", "zero_points", "qweight", "qzeros"))
            assert out[n]["qweight"].device.type == "cpu" and out[n]["qweight"].is_pinned() == pin


def test_model_level_search_failure_modes(native_lib, cuda_device, monkeypatch):
    """quantize_model(activations=...): (1) keep_unpacked reaches the tensors WITHOUT activations too; (2) when the
    streamed search fails as a whole, every tensor is still searched (one by one) instead of silently losing its AWQ
    scaling; (3) a pipelined section that fails is drained and its tensors take the per-tensor path"""
    from awq_quantizer.quantization import AWQQuantizer
    from awq_quantizer.quantization import arena as A
    from awq_quantizer.quantization import stream as S
    import awq_quantizer.quantization.search as SE
    qz = AWQQuantizer(bits=4, group_size=128, symmetric=False, device="cuda:0", logger_level="CRITICAL", n_grid=6)
    X = datagen.activations(96, 1024, "bf16", 9)
    tensors = {"a": datagen.weights((256, 1024), "bf16", 1), "b": datagen.weights((128, 1024), "bf16", 2),
               "emb": datagen.weights((64, 1024), "bf16", 3), "norm": datagen.weights((1024,), "bf16", 4)}
    acts = {"a": X, "b": X}
    good = qz.quantize_model(tensors, activations=acts, pack=True, keep_unpacked=True)
    assert list(good) == list(tensors)
    for n in tensors:                                                # (1) one schema for all tensors
        assert {"tensor_q", "zero_points", "qweight", "qzeros", "scales"} <= set(good[n]), n
    assert "awq_scale" in good["a"] and "awq_scale" not in good["emb"]
    want = O.pack_result(O.group_quant_vec(tensors["emb"], 4, 128, False, True))
    assert torch.equal(good["emb"]["qweight"], want["qweight"]) and torch.equal(good["emb"]["tensor_q"], want["tensor_q"])

    def boom(*a, **k):
        raise RuntimeError("injected failure of the streamed search")
    monkeypatch.setattr(SE, "quantize_model_with_search", boom)      # (2)
    fb = qz.quantize_model(tensors, activations=acts, pack=True, keep_unpacked=True)
    monkeypatch.undo()
    for n in ("a", "b"):
        assert int(fb[n]["best_idx"]) == int(good[n]["best_idx"]) and torch.equal(fb[n]["qweight"], good[n]["qweight"])
        assert torch.equal(fb[n]["awq_scale"], good[n]["awq_scale"])
    assert torch.equal(fb["emb"]["qweight"], good["emb"]["qweight"])

    calls = {"n": 0}
    real = A.quantize_arena

    def flaky(*a, **k):                                              # (3) the arena section fails once
        calls["n"] += 1
        if calls["n"] == 1:
            raise RuntimeError("injected failure of the arena pipeline")
        return real(*a, **k)
    monkeypatch.setattr(A, "quantize_arena", flaky)
    plain = {n: t for n, t in tensors.items()}
    res = qz.quantize_model(plain, pack=True)
    monkeypatch.undo()
    assert calls["n"] >= 1 and list(sorted(res)) == sorted(plain)
    for n, t in plain.items():
        w = O.pack_result(O.group_quant_vec(t, 4, 128, False, True))
        assert torch.equal(res[n]["qweight"], w["qweight"]) and torch.equal(res[n]["qzeros"], w["qzeros"]), n


def test_search_from_worker_threads(native_lib, cuda_device):
    """the reference calls quantize() from a ThreadPoolExecutor (main.py:609-621): concurrent searches (each with its
    own workspace; the cooperative launches serialise on the device) give the results of sequential calls"""
    from concurrent.futures import ThreadPoolExecutor
    from awq_quantizer.quantization import AWQQuantizer
    qz = AWQQuantizer(bits=4, group_size=128, symmetric=False, device="cuda:0", logger_level="ERROR", n_grid=8)
    jobs = [(datagen.weights((256 + 64 * i, 1024), "bf16", 30 + i), datagen.activations(128, 1024, "bf16", 40 + (i % 2)))
            for i in range(6)]
    want = [qz.quantize(w, activations=x, pack=True) for w, x in jobs]
    with ThreadPoolExecutor(max_workers=4) as ex:
        got = list(ex.map(lambda wx: qz.quantize(wx[0], activations=wx[1], pack=True), jobs))
    for g, w in zip(got, want):
        assert int(g["best_idx"]) == int(w["best_idx"])
        assert torch.equal(g["qweight"], w["qweight"]) and torch.equal(g["qzeros"], w["qzeros"])
        assert torch.allclose(g["search_err"], w["search_err"], rtol=1e-9, atol=0)

"""Comparison helpers shared by the tests."""
import torch


def canon(t: torch.Tensor) -> torch.Tensor:
    """fp16 -> int16 bit pattern with every NaN mapped to 0x7E00 (NaN sign / payload are not part of
    the parity contract); fp32 -> NaN mapped to one pattern; other dtypes unchanged."""
    t = t.detach().cpu()
    if t.dtype == torch.float16:
        bits = t.contiguous().view(torch.int16).clone()
        bits[torch.isnan(t)] = 0x7E00
        return bits
    if t.dtype == torch.float32:
        bits = t.contiguous().view(torch.int32).clone()
        bits[torch.isnan(t)] = 0x7FC00000
        return bits
    return t


def assert_same(a: torch.Tensor, b: torch.Tensor, what: str = ""):
    assert a.dtype == b.dtype, f"{what}: dtype {a.dtype} != {b.dtype}"
    assert tuple(a.shape) == tuple(b.shape), f"{what}: shape {tuple(a.shape)} != {tuple(b.shape)}"
    ca, cb = canon(a), canon(b)
    if not torch.equal(ca, cb):
        bad = (ca != cb).nonzero()
        first = tuple(bad[0].tolist()) if bad.numel() else ()
        raise AssertionError(f"{what}: {bad.shape[0]} of {ca.numel()} elements differ; first at {first}: "
                             f"{a[first] if first else a} vs {b[first] if first else b}")


def assert_quant_equal(got: dict, want: dict, what: str = "", keys=("tensor_q", "scales", "zero_points")):
    for k in keys:
        assert_same(got[k], want[k], f"{what}/{k}")

"""Case tables for the golden fixtures (shared by make_golden.py and the tests)."""
from __future__ import annotations

import itertools

import torch

from tests import datagen

# (shape, offset): ragged K with zero padding, 1-D, 3-D, numel<g bypass, 0-d.
SMALL_SHAPES = [
    ((8, 300), 0.3),
    ((300,), 0.3),
    ((8, 3, 100), 0.3),
    ((5, 129), -0.2),
    ((4, 1000), 0.0),
    ((130,), 0.3),
    ((16, 512), 0.0),
    ((3, 1024), 0.0),
    ((10, 10), 0.0),      # numel < g: per-channel scales [C]       (awq.py:297-300)
    ((50,), 0.1),         # numel < g, 1-D: 0-d scale
    ((), 0.0),            # 0-d tensor
    ((2, 3, 4, 5), 0.0),  # 4-D, numel = 120 < 128
]


def small_cases():
    """Grid over shapes x dtype x symmetric x bits x group size (+ per_channel
    False on the bypass shapes).  Yields dicts."""
    for (shape, off), dt, sym, bits, g in itertools.product(
            SMALL_SHAPES, ("bf16", "fp16", "fp32"), (False, True), (4, 8), (32, 64, 128)):
        yield dict(shape=shape, offset=off, dtype=dt, symmetric=sym, bits=bits, group_size=g,
                   per_channel=True, std=0.02)
    # per_channel=False only changes the numel < g bypass
    for (shape, off), dt, sym in itertools.product(
            [((10, 10), 0.0), ((2, 3, 4, 5), 0.0)], ("bf16", "fp32"), (False, True)):
        yield dict(shape=shape, offset=off, dtype=dt, symmetric=sym, bits=4, group_size=128,
                   per_channel=False, std=0.02)
    # odd group sizes (generic kernel path) and fp64 input
    for g, dt, sym in itertools.product((1, 7, 100, 256, 1000), ("bf16", "fp32"), (False, True)):
        yield dict(shape=(6, 520), offset=0.05, dtype=dt, symmetric=sym, bits=4, group_size=g,
                   per_channel=True, std=0.02)
    for sym, g in itertools.product((False, True), (64, 128)):
        yield dict(shape=(4, 300), offset=0.1, dtype="fp64", symmetric=sym, bits=4, group_size=g,
                   per_channel=True, std=0.02)
    # large magnitudes (fp16 scale overflow -> inf) and tiny magnitudes (fp16 scale flush)
    for std, dt, sym in itertools.product((3.0e4, 1.0e-7, 1.0), ("bf16", "fp16", "fp32"), (False, True)):
        yield dict(shape=(4, 256), offset=0.0, dtype=dt, symmetric=sym, bits=4, group_size=128,
                   per_channel=True, std=std)


def case_key(c) -> str:
    shp = "x".join(str(s) for s in c["shape"]) or "scalar"
    return (f"{shp}_{c['dtype']}_{'sym' if c['symmetric'] else 'asym'}_b{c['bits']}_g{c['group_size']}"
            f"_{'pc' if c['per_channel'] else 'pt'}_o{c['offset']}_s{c['std']}")


def case_input(c) -> torch.Tensor:
    seed = datagen.seed_of("small", c["shape"], c["dtype"], c["offset"], c["std"])
    return datagen.weights(c["shape"], c["dtype"], seed, std=c["std"], offset=c["offset"])


def special_inputs():
    """Hand-built tensors for the non-finite / degenerate branches (all [R,256],
    g=128).  Returned as fp32 masters; cast per dtype by the caller."""
    out = {}
    base = datagen.weights((4, 256), "fp32", datagen.seed_of("special"), std=0.02)
    t = base.clone(); t[0, 5] = float("nan"); t[1, 130] = float("inf"); t[2, 7] = float("-inf")
    out["nonfinite"] = t
    t = base.clone(); t[0, :128] = 0.0; t[1, 128:] = 0.0; t[2, :] = 0.0
    out["zero_groups"] = t
    t = base.clone(); t[0, :128] = 0.5; t[1, 128:] = -0.5; t[2, :128] = 1e-30; t[3, :] = 7.0
    out["constant_groups"] = t
    t = base.clone().abs() + 0.01                       # all positive -> zp clamps to 0
    out["all_positive"] = t
    t = -(base.clone().abs()) - 0.01                    # all negative -> zp clamps to qmax
    out["all_negative"] = t
    t = base.clone(); t[0, :128] *= 1e-38; t[1, :128] *= 1e-42; t[2, 128:] *= 1e30
    out["extreme_exponents"] = t
    t = base.clone(); t[:, ::2] = 0.0; t[0, 1] = -0.0
    out["sparse_signed_zero"] = t
    # exact rounding ties: values k + 0.5 times a power-of-two scale
    t = torch.zeros(4, 256)
    t[:, :128] = (torch.arange(128) % 16).float() * 0.25 + 0.125
    t[:, 0] = 0.0; t[:, 1] = 3.75
    t[:, 128:] = ((torch.arange(128) % 31).float() - 15.0) * 0.5
    out["ties"] = t
    return out


# Medium cases: digests only (inputs regenerated from the seed).  config-0 shapes
# from the reference's test_quantization.py:54-63,132,136-145.
MEDIUM_CASES = [
    dict(name="cfg0_ffn", shape=(768, 3072), src="bf16", convert_fp16=True, symmetric=True),
    dict(name="cfg0_conv", shape=(768, 3, 768), src="bf16", convert_fp16=True, symmetric=True),
    dict(name="cli_bf16_asym", shape=(256, 3072), src="bf16", convert_fp16=False, symmetric=False),
    dict(name="fp32_asym", shape=(256, 3072), src="fp32", convert_fp16=False, symmetric=False),
]


def medium_input(c) -> torch.Tensor:
    return datagen.weights(c["shape"], c["src"], datagen.seed_of("medium", c["name"]), std=0.02)

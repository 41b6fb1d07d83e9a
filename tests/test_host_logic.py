"""CPU: the Python mirror of the reference interface -- constructor / validation / layout rules /
loud failure without a GPU (no CPU fallback)."""
import pytest
import torch

from awq_quantizer.quantization import AWQQuantizer
from awq_quantizer.quantization.awq import AWQQuantizer as AWQQuantizer2


def mk(**kw):
    kw.setdefault("logger_level", "ERROR")
    return AWQQuantizer(**kw)


def test_same_import_paths_as_reference():
    assert AWQQuantizer is AWQQuantizer2


def test_defaults_match_reference_signature():
    q = mk()
    assert (q.bits, q.group_size, q.symmetric, q.zero_point, q.percentile, q.scale_method, q.per_channel) == \
        (4, 128, True, "minmax", 0.99, "mse", True)
    assert (q.qmin, q.qmax) == (-8, 7)
    assert (mk(symmetric=False).qmin, mk(symmetric=False).qmax) == (0, 15)
    assert (mk(bits=8).qmin, mk(bits=8).qmax) == (-128, 127)
    assert (mk(bits=8, symmetric=False).qmin, mk(bits=8, symmetric=False).qmax) == (0, 255)


@pytest.mark.parametrize("kw,msg", [
    (dict(bits=3), "Unsupported bit width: 3. Supported: 4, 8."),
    (dict(group_size=0), "Group size must be a positive integer: 0"),
    (dict(group_size=-4), "Group size must be a positive integer: -4"),
    (dict(group_size=64.0), "Group size must be a positive integer: 64.0"),
    (dict(zero_point="foo"), "Unsupported zero point calibration method: foo"),
    (dict(zero_point="percentile", percentile=1.5), "Percentile must be in range (0, 1): 1.5"),
    (dict(scale_method="abs"), "Unsupported scale calibration method: abs"),
])
def test_validation_messages(kw, msg):
    with pytest.raises(ValueError) as e:
        mk(**kw)
    assert str(e.value) == msg


def test_type_errors_precede_device_use():
    q = mk(device="cuda")
    with pytest.raises(ValueError, match="Expected torch.Tensor"):
        q.quantize([1.0, 2.0])
    with pytest.raises(ValueError, match="Expected floating point tensor"):
        q.quantize(torch.zeros(100, 100, dtype=torch.int32))   # test_quantization.py:63


def test_no_cpu_fallback():
    q = mk(device="cpu")
    with pytest.raises(RuntimeError, match="no CPU"):
        q.quantize(torch.zeros(4, 128))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            mk(device="cuda").quantize(torch.zeros(4, 128))
        with pytest.raises(RuntimeError):
            mk(device="cuda").dequantize({"tensor_q": torch.zeros(1, 128, dtype=torch.int32),
                                          "scales": torch.ones(1, 1, dtype=torch.float16),
                                          "zero_points": torch.zeros(1, 1, dtype=torch.int32),
                                          "group_size": torch.tensor(128)})


def test_quantize_model_swallows_per_tensor_errors():
    # awq.py:453-455: errors are logged and the tensor is skipped
    q = mk(device="cpu")
    assert q.quantize_model({"a": torch.zeros(4, 128), "b": torch.zeros(3, dtype=torch.int64)}) == {}


@pytest.mark.parametrize("shape,pc,expect", [
    ((768, 3072), True, (768, 3072, 128, 24, (768, 24))),
    ((768, 3, 768), True, (768, 2304, 128, 18, (768, 18))),
    ((8, 300), True, (8, 300, 128, 3, (8, 3))),
    ((300,), True, (1, 300, 128, 3, (1, 3))),
    ((10, 10), True, (10, 10, 10, 1, (10,))),          # bypass, per-channel -> scales [C]
    ((10, 10), False, (1, 100, 100, 1, ())),           # bypass, per-tensor -> 0-d scale
    ((50,), True, (1, 50, 50, 1, ())),
    ((), True, (1, 1, 1, 1, ())),
    ((2, 3, 4, 5), True, (2, 60, 60, 1, (2,))),
])
def test_layout_rules(shape, pc, expect):
    q = mk(per_channel=pc)
    assert q._layout(torch.zeros(shape)) == expect


def test_model_shape_inventories_match_survey():
    """parameter counts of the BASELINE.json model shapes"""
    from awq_quantizer import model_shapes as M
    assert M.total_params(M.workload("opt-125m")) == 125_237_760
    assert abs(M.total_params(M.workload("opt-350m")) / 1e9 - 0.331) < 0.001
    assert abs(M.total_params(M.workload("llama3-8b")) / 1e9 - 8.030) < 0.001
    assert abs(M.total_params(M.workload("llama3-70b")) / 1e9 - 70.55) < 0.01
    assert [s for _, s, _ in M.workload("test_quantization")] == [(768, 3072), (768, 3, 768), (10, 10)]
    assert M.workload("micro-8192x28672")[0][1] == (8192, 28672)
    lin = [s for _, s, ck in M.workload("llama3-8b") if ck is not None]
    assert len(lin) == 32 * 7 and sum(a * b for a, b in lin) == 6_979_321_856     # 6.979 B linear params


def test_numa_binding_is_best_effort():
    from awq_quantizer import parallel
    assert parallel.bind_to_gpu_numa(0) in (None, 0, 1, 2, 3, 4, 5, 6, 7)        # never raises (no GPU here -> None)
    assert parallel.rank_info() == (0, 1, 0)
    assert parallel.gather_metadata({"rank": 0}) == [{"rank": 0}]


def test_product_package_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under awq-converter_b200/ may import, load or execute it"""
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "awq-converter_b200")
    bad = []
    for d, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(d, f), errors="ignore").read()
                # (comments may cite oracle/awq_oracle.py as the definition of the unpinned parts)
                if re.search(r"^\s*(from|import)\s+oracle\b|libawq_oracle|import_module\(.oracle", text, flags=re.M):
                    bad.append(os.path.relpath(os.path.join(d, f), root))
    assert not bad, bad


def test_arena_layout_and_short_row_classes():
    """host-side layout rules of the pipelines (no device): tile-aligned virtual arena, eligibility classes"""
    from awq_quantizer.quantization.arena import TILE, HostArena, arena_eligible, pipe_eligible, short_row_len
    bf, f32 = torch.bfloat16, torch.float32
    assert arena_eligible((4096, 1024), bf, 128, 4) and arena_eligible((1024,), bf, 128, 4)
    assert not arena_eligible((1024, 512), bf, 128, 4) and pipe_eligible((1024, 512), bf, 128, 4)     # 4 groups per row
    assert not pipe_eligible((7, 300), bf, 128, 4) and not pipe_eligible((10, 10), bf, 128, 4)
    assert not pipe_eligible((8, 1024), torch.float64, 128, 4) and not pipe_eligible((8, 1024), bf, 96, 4)
    assert short_row_len((1024, 512), bf, 128, 4) == 512 and short_row_len((9, 128), bf, 128, 4) == 128
    assert short_row_len((40, 256), f32, 128, 4) == 256 and short_row_len((8, 64), bf, 32, 4) == 64
    assert short_row_len((8, 1536), bf, 128, 4) == 0            # 12 groups: neither a whole word nor a short row
    assert short_row_len((8, 384), bf, 128, 4) == 0             # 3 groups do not divide a word
    assert short_row_len((8, 1024), bf, 128, 4) == 0            # whole words: the flat class
    assert short_row_len((8, 512), bf, 128, 8) == 0 and short_row_len((8, 256), bf, 128, 8) == 256    # int8: 4 per word
    tensors = {"a": torch.zeros(3, 1024, dtype=bf), "b": torch.zeros(1024, dtype=bf), "h": torch.zeros(5, 2048, dtype=torch.float16)}
    arena = HostArena.for_tensors(tensors)
    assert not arena.buffers and arena.sizes == {bf: 2 * TILE, torch.float16: 2 * TILE}
    assert arena.layout[bf] == [("a", 0, 3072), ("b", TILE, 1024)] and arena.layout[torch.float16] == [("h", 0, 10240)]
    real = HostArena.from_tensors({k: v + 1 for k, v in tensors.items()}, pin=False)
    assert real.views["b"].shape == (1024,) and float(real.buffers[bf][TILE]) == 1.0
    assert float(real.buffers[bf][3072]) == 0.0 and float(real.buffers[bf][TILE - 1]) == 0.0          # padding is zeroed
    assert real.payload_bytes() == (3072 + 1024 + 10240) * 2


def test_stream_wave_planning():
    """host logic of the streamed model-level AWQ: waves in input order, oversize tensors alone, ring-slot sizing"""
    from awq_quantizer.quantization import stream
    t = {"a": torch.zeros(64, 256, dtype=torch.bfloat16), "b": torch.zeros(64, 256, dtype=torch.bfloat16),
         "big": torch.zeros(512, 1024, dtype=torch.bfloat16), "c": torch.zeros(16, 256, dtype=torch.float16)}
    assert stream.plan_waves(list(t), t, 70_000) == [["a", "b"], ["big"], ["c"]]
    assert stream.plan_waves(list(t), t, 1) == [["a"], ["b"], ["big"], ["c"]]
    assert stream.plan_waves(list(t), t, 1 << 30) == [list(t)]
    assert stream.plan_waves([], t, 1) == []
    kw = dict(group_size=128, bits=4, n_grid=20)
    base = stream.result_bytes(t["a"], pack=False, keep_unpacked=False, **kw)
    packed = stream.result_bytes(t["a"], pack=True, keep_unpacked=False, **kw)
    both = stream.result_bytes(t["a"], pack=True, keep_unpacked=True, **kw)
    assert packed - base == 64 * 32 * 4 + 256 and both - packed == 64 * 256 * 4      # qweight + one padded qzeros slot; int32 codes
    assert base >= 64 * 2 * 2 + 64 * 2 * 4 + 20 * 8 + 4 + 256 * 4 and base % 256 == 0


def test_activation_buffer_plan():
    """calibration activations share a few device buffers per shape class: occupants of one buffer never overlap
    (with the uploader's lookahead), and every tensor knows which wave released its buffer"""
    from awq_quantizer.quantization.stream import plan_activation_buffers
    xs = [torch.zeros(4, 8) for _ in range(6)] + [torch.zeros(4, 16)]
    acts = {f"t{i}": xs[i // 2] for i in range(12)}
    acts["wide"] = xs[6]
    waves = [[f"t{i}"] for i in range(12)] + [["wide"]]
    plan = plan_activation_buffers(waves, acts, torch.device("cpu"), lookahead=2)
    assert set(plan) == {id(x) for x in xs}
    first = {id(xs[i]): 2 * i for i in range(6)}
    last = {id(xs[i]): 2 * i + 1 for i in range(6)}
    by_buf = {}
    for k, (buf, prev_last) in plan.items():
        assert tuple(buf.shape) == tuple([x for x in xs if id(x) == k][0].shape)
        by_buf.setdefault(id(buf), []).append((k, prev_last))
    assert len(by_buf) == 2 + 1                                  # 2 rotating 4x8 buffers at lookahead 2, one 4x16
    for occupants in by_buf.values():
        occupants = [o for o in occupants if o[0] in first]
        occupants.sort(key=lambda o: first[o[0]])
        for (ka, _), (kb, prev) in zip(occupants, occupants[1:]):
            assert prev == last[ka] and last[ka] < first[kb] - 2

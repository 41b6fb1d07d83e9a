"""CPU: the callers either side of the hot path -- CLI flags, tensor selection, LPT partitioning,
chunked save (.pt and the fixed flat-key safetensors layout), loader contract, YAML config, and the
world-size-2 (gloo) shard + metadata-gather logic."""
import json
import os
import subprocess
import sys

import pytest
import torch

from awq_quantizer import main as cli
from awq_quantizer import model_shapes as M
from awq_quantizer import parallel
from awq_quantizer.model_loading import SafetensorsLoader, load_model_from_hub, load_model_from_path, verify_file_hash
from awq_quantizer.utils.config import load_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_FLAGS = ["--model_id", "--output_dir", "--bits", "--group_size", "--symmetric", "--zero_point", "--percentile",
             "--scale_method", "--per_channel", "--device", "--num_workers", "--max_memory", "--multi_gpu",
             "--batch_size", "--prefetch_factor", "--memory_efficient", "--log_level", "--log_file",
             "--save_safetensors", "--chunk_size"]      # main.py:32-157


def test_cli_flags_and_defaults_match_reference():
    a = cli.parse_args(["--model_id", "m", "--output_dir", "o"])
    assert (a.bits, a.group_size, a.symmetric, a.zero_point, a.percentile, a.scale_method, a.per_channel) == \
        (4, 128, False, "minmax", 0.99, "mse", False)
    assert (a.num_workers, a.max_memory, a.multi_gpu, a.batch_size, a.prefetch_factor, a.chunk_size) == \
        (4, 0.8, False, 10, 2, 10)
    assert a.save_safetensors is False and a.log_level == "INFO" and a.pack is False and a.arith == "native"
    assert a.packed_only is False
    helptext = subprocess.run([sys.executable, "-c",
                               "import sys; sys.path.insert(0, %r); from awq_quantizer.main import parse_args; "
                               "parse_args(['--help'])" % os.path.join(ROOT, "awq-converter_b200")],
                              capture_output=True, text=True).stdout
    for f in REF_FLAGS:
        assert f in helptext, f
    with pytest.raises(SystemExit):
        cli.parse_args(["--model_id", "m", "--output_dir", "o", "--bits", "3"])


def test_tensor_selection_rules():
    t = {"big": torch.zeros(64, 256), "small": torch.zeros(10, 10), "ints": torch.zeros(200, 2, dtype=torch.int32),
         "empty": torch.zeros(0, 4), "mid": torch.zeros(300), "notensor": 3}
    batches = cli.prepare_tensors_for_quantization(t, "cpu", batch_size=1)
    names = [n for b in batches for n in b]
    assert names == ["big", "mid"]                      # largest first; numel < 128, non-float, empty dropped
    assert all(len(b) == 1 for b in batches)


def test_partition_is_lpt_and_complete():
    specs = M.workload("opt-125m")
    tensors = {n: torch.empty(s, dtype=torch.bfloat16, device="meta") for n, s, _ in specs}
    parts = cli.partition_tensors(tensors, 4)
    assert sorted(n for p in parts for n in p) == sorted(tensors)
    loads = [sum(t.numel() * 2 for t in p.values()) for p in parts]
    biggest = max(t.numel() * 2 for t in tensors.values())
    assert max(loads) - min(loads) <= biggest           # LPT bound
    assert cli.partition_tensors(tensors, 1) == [tensors]
    # deterministic
    assert [list(p) for p in cli.partition_tensors(tensors, 4)] == [list(p) for p in parts]


def fake_qdict(shape, pack=False):
    C, K = shape
    d = {"tensor_q": torch.zeros(shape, dtype=torch.int32), "scales": torch.ones((C, K // 128), dtype=torch.float16),
         "zero_points": torch.zeros((C, K // 128), dtype=torch.int32), "bits": torch.tensor(4, dtype=torch.int32),
         "group_size": torch.tensor(128, dtype=torch.int32), "symmetric": torch.tensor(False)}
    if pack:
        d["qweight"] = torch.zeros((C, K // 8), dtype=torch.int32)
    return d


def test_save_chunks_pt_and_safetensors(tmp_path):
    q = {f"t{i}": fake_qdict((4, 256), pack=(i % 2 == 0)) for i in range(5)}
    meta = cli.save_model_in_chunks(q, str(tmp_path / "pt"), chunk_size=2)
    md = json.load(open(tmp_path / "pt" / "metadata.json"))
    assert md["num_chunks"] == 3 and md["num_tensors"] == 5 and md["format"] == "pytorch"
    assert md["quantization_params"] == {"bits": 4, "group_size": 128, "symmetric": False}
    assert md["tensor_to_chunk"]["t4"] == 2 and meta["files"][0] == "model_chunk_0000.pt"
    back = torch.load(tmp_path / "pt" / "model_chunk_0001.pt")
    assert sorted(back) == ["t2", "t3"] and torch.equal(back["t2"]["scales"], q["t2"]["scales"])
    # the reference's --save_safetensors path raises on nested dicts (main.py:478-490); here: flat keys
    cli.save_model_in_chunks(q, str(tmp_path / "st"), chunk_size=10, use_safetensors=True)
    from safetensors.torch import load_file
    flat = load_file(str(tmp_path / "st" / "model_chunk_0000.safetensors"))
    for suffix in ("q", "scales", "zero_points", "bits", "group_size", "symmetric"):     # test_quantization.py:182-189
        assert f"t1.{suffix}" in flat
    assert flat["t1.q"].dtype == torch.int32 and flat["t1.scales"].dtype == torch.float16
    assert "t0.qweight" in flat and "t1.qweight" not in flat
    # results that are views of one model-sized array (the pipeline's layout) must not drag it into every chunk
    big = torch.arange(1 << 20, dtype=torch.int32)
    v = {f"v{i}": dict(fake_qdict((4, 256)), tensor_q=big[i * 1024:(i + 1) * 1024].view(4, 256)) for i in range(4)}
    cli.save_model_in_chunks(v, str(tmp_path / "views"), chunk_size=1)
    import os
    assert all(os.path.getsize(tmp_path / "views" / f"model_chunk_{i:04d}.pt") < 64 << 10 for i in range(4))
    back = torch.load(tmp_path / "views" / "model_chunk_0002.pt")
    assert torch.equal(back["v2"]["tensor_q"], v["v2"]["tensor_q"])


def test_loader_contract_and_arena(tmp_path):
    from safetensors.torch import save_file
    a = {"w1": torch.randn(8, 256).to(torch.bfloat16), "b1": torch.randn(256).to(torch.bfloat16)}
    b = {"w2": torch.randn(4, 1024).to(torch.float16), "odd": torch.randn(3, 100), "ids": torch.arange(10)}
    save_file(a, str(tmp_path / "model-00001-of-00002.safetensors"))
    save_file(b, str(tmp_path / "model-00002-of-00002.safetensors"))
    save_file({"w1": torch.zeros(1)}, str(tmp_path / "consolidated.safetensors"))      # ignored when shards exist
    for ld in (load_model_from_path(str(tmp_path), logger_level="ERROR"),
               load_model_from_hub(str(tmp_path), logger_level="ERROR")):               # local dir as model id
        t = ld.load_tensors()
        assert sorted(t) == ["b1", "ids", "odd", "w1", "w2"]
        assert t["w1"].dtype == torch.bfloat16 and t["w1"].device.type == "cpu" and torch.equal(t["w1"], a["w1"])
    arena, rest = ld.load_arena(group_size=128, bits=4)
    assert sorted(arena.views) == ["w2"] and sorted(rest) == ["b1", "ids", "odd", "w1"]   # G % 8 == 0 only for w2
    assert torch.equal(arena.views["w2"], b["w2"])
    with pytest.raises(ValueError, match="No safetensor files found"):
        SafetensorsLoader(str(tmp_path / "nothing"))
    h = verify_file_hash(str(tmp_path / "consolidated.safetensors"))
    assert len(h) == 64
    with pytest.raises(ValueError):
        verify_file_hash(str(tmp_path / "consolidated.safetensors"), "00")


def test_yaml_config_schema(tmp_path):
    c = load_config()
    assert c.get("quantization.bits") == 4 and c.get("quantization.group_size") == 128
    assert c.get("quantization.symmetric") is True and c.get("quantization.skip_layers") == []
    assert c.get("hardware.device") == "cuda" and c.get("output.safetensors") is True and c["logging.level"] == "INFO"
    p = tmp_path / "u.yaml"
    p.write_text("quantization:\n  bits: 8\n  skip_layers: [lm_head]\nmodel:\n  path: facebook/opt-350m\n")
    c = load_config(str(p))
    assert c.get("quantization.bits") == 8 and c.get("quantization.group_size") == 128
    assert c.get("model.from_hub") is True and c.get("model.hub_model_id") == "facebook/opt-350m"
    c.set("a.b.c", 5)
    assert c.get("a.b.c") == 5 and c.get("nope", "d") == "d"


def test_cli_refuses_cpu(tmp_path):
    from safetensors.torch import save_file
    save_file({"w": torch.randn(4, 256)}, str(tmp_path / "m.safetensors"))
    rc = cli.main(["--model_id", str(tmp_path), "--output_dir", str(tmp_path / "out"), "--device", "cpu",
                   "--log_level", "ERROR"])
    assert rc == 1                                       # no CPU fallback; the CLI reports and returns 1


def test_merge_chunk_maps():
    metas = [{"rank": 1, "num_chunks": 2, "tensor_to_chunk": {"c": 0, "d": 1}, "files": ["r1a", "r1b"]},
             {"rank": 0, "num_chunks": 1, "tensor_to_chunk": {"a": 0, "b": 0}, "files": ["r0a"]}]
    m = parallel.merge_chunk_maps(metas)
    assert m["num_chunks"] == 3 and m["tensor_to_chunk"] == {"a": 0, "b": 0, "c": 1, "d": 2}
    assert m["files"] == ["r0a", "r1a", "r1b"]
    with pytest.raises(ValueError):
        parallel.merge_chunk_maps([metas[0], {**metas[0], "rank": 2}])


WORKER = r"""
import os, sys, json
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], 'awq-converter_b200'))
import torch
from awq_quantizer import parallel, model_shapes as M
rank, world = parallel.init_distributed('gloo')
specs = M.workload('opt-125m')
items = [(n, M.numel(s) * 2) for n, s, _ in specs]
mine = parallel.shard_for_rank(items, world, rank)
meta = {'rank': rank, 'num_chunks': (len(mine) + 9) // 10, 'tensor_to_chunk': {n: i // 10 for i, n in enumerate(mine)},
        'files': [f'rank{rank}_model_chunk_{c:04d}.pt' for c in range((len(mine) + 9) // 10)],
        'bytes': sum(dict(items)[n] for n in mine)}
metas = parallel.gather_metadata(meta)
if rank == 0:
    merged = parallel.merge_chunk_maps(metas)
    json.dump({'merged_tensors': merged['num_tensors'], 'num_chunks': merged['num_chunks'], 'all': len(items),
               'bytes': [m['bytes'] for m in sorted(metas, key=lambda m: m['rank'])],
               'max_item': max(b for _, b in items)}, open(sys.argv[2], 'w'))
torch.distributed.destroy_process_group()
"""


def test_two_rank_gloo_shard_and_gather(tmp_path):
    """N>1 host path on CPU: world_size 2 over gloo, rank-local LPT shards, one all_gather_object"""
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = tmp_path / "out.json"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script), ROOT, str(out)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.load(open(out))
    assert res["merged_tensors"] == res["all"] == 194          # every tensor exactly once
    assert abs(res["bytes"][0] - res["bytes"][1]) <= res["max_item"]


def test_header_index_named_loading_and_selection(tmp_path):
    """the orchestrator partitions on the header index and every rank reads only its own tensors"""
    from safetensors.torch import save_file
    a = {"w1": torch.randn(8, 256).to(torch.bfloat16), "b1": torch.randn(64).to(torch.bfloat16)}
    b = {"w2": torch.randn(4, 1024).to(torch.float16), "ids": torch.arange(200), "lm_head.weight": torch.randn(2, 128),
         "empty": torch.zeros(0, 4)}
    save_file(a, str(tmp_path / "model-00001-of-00002.safetensors"))
    save_file(b, str(tmp_path / "model-00002-of-00002.safetensors"))
    ld = load_model_from_path(str(tmp_path), logger_level="ERROR")
    idx = ld.index()
    assert sorted(idx) == ["b1", "empty", "ids", "lm_head.weight", "w1", "w2"]
    assert idx["w1"].shape == (8, 256) and idx["w1"].dtype == torch.bfloat16 and idx["w1"].nbytes == 4096
    assert idx["w2"].path.endswith("00002-of-00002.safetensors") and idx["ids"].dtype == torch.int64
    quant, passed = cli.select_tensors(idx, skip_layers=["lm_head"])
    assert quant == ["w2", "w1"]                                     # largest first (main.py:256)
    assert sorted(passed) == ["b1", "empty", "ids", "lm_head.weight"]  # small / non-float / empty / skip_layers
    got = ld.load_tensors(names=["w2", "b1"])
    assert list(got) == ["b1", "w2"] and torch.equal(got["w2"], b["w2"]) and torch.equal(got["b1"], a["b1"])
    with pytest.raises(KeyError):
        ld.load_tensors(names=["nope"])
    # two ranks: disjoint, complete, and each asks the loader for its own names only
    costs = [(n, idx[n].nbytes) for n in quant + passed]
    r0, r1 = (set(parallel.shard_for_rank(costs, 2, r)) for r in (0, 1))
    assert not (r0 & r1) and (r0 | r1) == set(idx)


@pytest.mark.timeout(300)
def test_cli_two_ranks_fail_together_without_hanging(tmp_path):
    """a rank that cannot do its work still reaches the metadata gather (here: both refuse the CPU); the job ends
    with exit code 1 instead of waiting for the collective's timeout, and the process group is destroyed"""
    from safetensors.torch import save_file
    save_file({"w": torch.randn(4, 256)}, str(tmp_path / "m.safetensors"))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", PYTHONPATH=os.path.join(ROOT, "awq-converter_b200"))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", "-m", "awq_quantizer.main",
                        "--model_id", str(tmp_path), "--output_dir", str(tmp_path / "out"), "--log_level", "ERROR"],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode != 0
    assert "no CPU execution path" in (r.stderr + r.stdout)
    assert not os.path.exists(tmp_path / "out" / "metadata.json") or \
        json.load(open(tmp_path / "out" / "metadata.json"))["num_tensors"] == 0


def test_native_safetensors_writer_round_trip(tmp_path):
    """write_safetensors (straight from the tensors' memory, no serialisation buffer) produces files the stock
    safetensors reader accepts, for every dtype / rank the result dicts hold, views of larger arrays included"""
    from safetensors.torch import load_file
    from safetensors import safe_open
    big = torch.arange(4096, dtype=torch.int32)
    flat = {"a.q": big[128:128 + 64].view(4, 16), "a.scales": torch.randn(4, 2).to(torch.float16),
            "a.zero_points": torch.arange(8, dtype=torch.int32).view(4, 2), "a.bits": torch.tensor(4, dtype=torch.int32).reshape(1),
            "a.symmetric": torch.tensor(True).reshape(1), "b.w": torch.randn(3, 5).to(torch.bfloat16),
            "b.f": torch.randn(7), "b.d": torch.randn(2, 2, dtype=torch.float64), "b.i64": torch.arange(3),
            "b.u8": torch.arange(5, dtype=torch.uint8), "b.i8": torch.arange(-2, 3, dtype=torch.int8),
            "b.empty": torch.empty((0, 4), dtype=torch.float16), "b.t": torch.arange(12, dtype=torch.int32).view(3, 4).t()}
    path = str(tmp_path / "x.safetensors")
    cli.write_safetensors(flat, path)
    back = load_file(path)
    assert sorted(back) == sorted(flat)
    for k, v in flat.items():
        assert back[k].dtype == v.dtype and back[k].shape == v.shape and torch.equal(back[k], v), k
    with safe_open(path, framework="pt") as f:                 # header is well formed for slicing readers too
        assert f.get_slice("a.q").get_shape() == [4, 16]
    with pytest.raises(ValueError):
        cli.write_safetensors({"c": torch.zeros(2, dtype=torch.complex64)}, str(tmp_path / "y.safetensors"))
    # many chunks -> several writer threads
    q = {f"t{i}": fake_qdict((4, 256), pack=True) for i in range(23)}
    meta = cli.save_model_in_chunks(q, str(tmp_path / "many"), chunk_size=2, use_safetensors=True)
    assert meta["num_chunks"] == 12 and meta["files"] == [f"model_chunk_{c:04d}.safetensors" for c in range(12)]
    got = load_file(str(tmp_path / "many" / "model_chunk_0011.safetensors"))
    assert torch.equal(got["t22.qweight"], q["t22"]["qweight"])


def test_calibration_file_aliases(tmp_path):
    """several weights may share one stored activation tensor through __metadata__ aliases: the index reports tokens
    for all of them and the loaded dict hands out the SAME tensor object (one upload, one scale grid)"""
    from safetensors.torch import save_file
    x = torch.randn(24, 64).to(torch.bfloat16)
    y = torch.randn(8, 32).to(torch.bfloat16)
    path = str(tmp_path / "calib.safetensors")
    save_file({"l0.q_proj.weight": x, "l0.down_proj.weight": y}, path,
              metadata={"alias.l0.k_proj.weight": "l0.q_proj.weight", "alias.l0.v_proj.weight": "l0.q_proj.weight",
                        "alias.dangling": "missing", "comment": "ignored"})
    idx = cli.read_calibration_index(path)
    assert idx == {"l0.q_proj.weight": 24, "l0.down_proj.weight": 8, "l0.k_proj.weight": 24, "l0.v_proj.weight": 24}
    calib = cli.load_calibration(path)
    assert sorted(calib) == sorted(idx)
    assert calib["l0.k_proj.weight"] is calib["l0.q_proj.weight"] and calib["l0.v_proj.weight"] is calib["l0.q_proj.weight"]
    assert torch.equal(calib["l0.q_proj.weight"], x) and torch.equal(calib["l0.down_proj.weight"], y)
    plain = str(tmp_path / "plain.safetensors")                 # files without metadata keep working
    save_file({"w": x}, plain)
    assert cli.read_calibration_index(plain) == {"w": 24} and list(cli.load_calibration(plain)) == ["w"]

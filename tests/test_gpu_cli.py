"""GPU: the CLI end to end on a tiny local model directory (the reference's offline path:
--model_id <dir>), .pt and safetensors outputs, against the oracle."""
import json
import os

import pytest
import torch

from oracle import awq_oracle as O
from tests import datagen
from tests.util import assert_quant_equal

pytestmark = pytest.mark.gpu


def make_model(tmp_path):
    from safetensors.torch import save_file
    t = {"layers.0.fc1.weight": datagen.weights((64, 256), "bf16", 1), "layers.0.fc1.bias": datagen.weights((256,), "bf16", 2),
         "layers.0.fc2.weight": datagen.weights((32, 512), "bf16", 3), "norm.weight": datagen.weights((100,), "bf16", 4),
         "ragged.weight": datagen.weights((6, 200), "fp32", 5), "position_ids": torch.arange(16)}
    d = tmp_path / "model"
    d.mkdir()
    save_file({k: v for k, v in list(t.items())[:3]}, str(d / "model-00001-of-00002.safetensors"))
    save_file({k: v for k, v in list(t.items())[3:]}, str(d / "model-00002-of-00002.safetensors"))
    return str(d), t


def test_cli_pt_output_matches_reference_layout(native_lib, cuda_device, tmp_path):
    from awq_quantizer import main as cli
    model, tensors = make_model(tmp_path)
    out = str(tmp_path / "out")
    rc = cli.main(["--model_id", model, "--output_dir", out, "--device", "cuda:0", "--chunk_size", "2",
                   "--log_level", "ERROR", "--num_workers", "2"])
    assert rc == 0
    md = json.load(open(os.path.join(out, "metadata.json")))
    expect = ["layers.0.fc1.weight", "layers.0.fc1.bias", "layers.0.fc2.weight", "ragged.weight"]   # numel >= 128, float
    assert sorted(md["tensor_to_chunk"]) == sorted(expect) and md["num_chunks"] == 2 and md["format"] == "pytorch"
    assert md["quantization_params"] == {"bits": 4, "group_size": 128, "symmetric": False}        # CLI default: asymmetric
    got = {}
    for c in range(md["num_chunks"]):
        got.update(torch.load(os.path.join(out, f"model_chunk_{c:04d}.pt")))
    for n in expect:
        assert_quant_equal(got[n], O.group_quant_vec(tensors[n], 4, 128, False, False), n)   # CLI default per_channel=False


def test_cli_safetensors_packed_symmetric(native_lib, cuda_device, tmp_path):
    from safetensors.torch import load_file
    from awq_quantizer import main as cli
    model, tensors = make_model(tmp_path)
    out = str(tmp_path / "out_st")
    rc = cli.main(["--model_id", model, "--output_dir", out, "--device", "cuda:0", "--save_safetensors", "--pack",
                   "--symmetric", "--per_channel", "--log_level", "ERROR"])
    assert rc == 0                                              # the reference returns 1 here (nested dicts, main.py:478-490)
    flat = load_file(os.path.join(out, "model_chunk_0000.safetensors"))
    for n in ("layers.0.fc1.weight", "layers.0.fc2.weight", "ragged.weight"):
        want = O.pack_result(O.group_quant_vec(tensors[n], 4, 128, True, True))
        assert torch.equal(flat[n + ".q"], want["tensor_q"]) and torch.equal(flat[n + ".qweight"], want["qweight"])
        assert torch.equal(flat[n + ".scales"].view(torch.int16), want["scales"].view(torch.int16))
        assert torch.equal(flat[n + ".zero_points"], want["zero_points"]) and torch.equal(flat[n + ".qzeros"], want["qzeros"])
        assert int(flat[n + ".bits"]) == 4 and int(flat[n + ".group_size"]) == 128 and bool(flat[n + ".symmetric"])


def test_cli_multi_gpu_flag_partitions_instead_of_repeating(native_lib, cuda_device, tmp_path):
    from awq_quantizer import main as cli
    model, tensors = make_model(tmp_path)
    out = str(tmp_path / "out_mg")
    rc = cli.main(["--model_id", model, "--output_dir", out, "--multi_gpu", "--log_level", "ERROR"])
    assert rc == 0
    md = json.load(open(os.path.join(out, "metadata.json")))
    assert md["num_tensors"] == 4


def test_cli_calibration_file_runs_search(native_lib, cuda_device, tmp_path):
    from safetensors.torch import save_file
    from awq_quantizer import main as cli
    from awq_quantizer.quantization import AWQQuantizer
    model, tensors = make_model(tmp_path)
    X = datagen.activations(96, 256, "bf16", 9)
    calib = str(tmp_path / "calib.safetensors")
    save_file({"layers.0.fc1.weight": X}, calib)
    out = str(tmp_path / "out_awq")
    rc = cli.main(["--model_id", model, "--output_dir", out, "--device", "cuda:0", "--calibration_file", calib,
                   "--n_grid", "8", "--pack", "--log_level", "ERROR", "--chunk_size", "100"])
    assert rc == 0
    got = torch.load(os.path.join(out, "model_chunk_0000.pt"))
    qz = AWQQuantizer(bits=4, group_size=128, symmetric=False, per_channel=False, device="cuda:0", logger_level="ERROR", n_grid=8)
    want = qz.quantize(tensors["layers.0.fc1.weight"], activations=X, pack=True)
    r = got["layers.0.fc1.weight"]
    assert int(r["best_idx"]) == int(want["best_idx"]) and torch.equal(r["qweight"], want["qweight"])
    assert torch.equal(r["awq_scale"], want["awq_scale"])
    assert "awq_scale" not in got["layers.0.fc2.weight"]
    # --packed_only: the same packed results without the int32 codes (searched and plain tensors alike)
    out2 = str(tmp_path / "out_awq_packed")
    rc = cli.main(["--model_id", model, "--output_dir", out2, "--device", "cuda:0", "--calibration_file", calib,
                   "--n_grid", "8", "--pack", "--packed_only", "--log_level", "ERROR", "--chunk_size", "100"])
    assert rc == 0
    got2 = torch.load(os.path.join(out2, "model_chunk_0000.pt"))
    assert set(got2) == set(got)
    for name, r2 in got2.items():
        assert "tensor_q" not in r2 and "tensor_q" in got[name]
        for k in ("qweight", "qzeros", "scales"):
            assert torch.equal(r2[k], got[name][k]), (name, k)
        if "zero_points" in r2:                    # (kept by the search path: small; the packed form is qzeros)
            assert torch.equal(r2["zero_points"], got[name]["zero_points"]), name


def test_cli_passthrough_of_unquantized_tensors(native_lib, cuda_device, tmp_path):
    """tensors the selection rules leave out (numel < 128, non-float) are not dropped as in the reference
    (main.py:243-253): bf16 goes through K3 (bf16 -> fp16, tensor_utils.py:10-22), the rest is copied unchanged"""
    from safetensors.torch import load_file
    from awq_quantizer import main as cli
    model, tensors = make_model(tmp_path)
    out = str(tmp_path / "out_pt")
    rc = cli.main(["--model_id", model, "--output_dir", out, "--device", "cuda:0", "--log_level", "ERROR"])
    assert rc == 0
    md = json.load(open(os.path.join(out, "metadata.json")))
    assert md["passthrough"] == {"norm.weight": "passthrough.safetensors", "position_ids": "passthrough.safetensors"}
    p = load_file(os.path.join(out, "passthrough.safetensors"))
    assert p["norm.weight"].dtype == torch.float16
    assert torch.equal(p["norm.weight"].view(torch.int16), O.bf16_to_fp16(tensors["norm.weight"]).view(torch.int16))
    assert torch.equal(p["position_ids"], tensors["position_ids"])


@pytest.mark.timeout(600)
def test_cli_under_torchrun_two_ranks(native_lib, cuda_device, tmp_path):
    """one process per rank, each loading and quantizing only its LPT share, one metadata gather (gloo here: the
    test box may have a single GPU, which both ranks then share -- NCCL refuses duplicate devices)"""
    import subprocess
    import sys
    model, tensors = make_model(tmp_path)
    out = str(tmp_path / "out_tr")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, AWQ_DIST_BACKEND="gloo", PYTHONPATH=os.path.join(root, "awq-converter_b200"))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29543", "-m", "awq_quantizer.main",
                        "--model_id", model, "--output_dir", out, "--pack", "--save_safetensors", "--log_level", "ERROR"],
                       capture_output=True, text=True, env=env, timeout=500)
    assert r.returncode == 0, r.stderr[-3000:]
    md = json.load(open(os.path.join(out, "metadata.json")))
    expect = ["layers.0.fc1.weight", "layers.0.fc1.bias", "layers.0.fc2.weight", "ragged.weight"]
    assert sorted(md["tensor_to_chunk"]) == sorted(expect) and md["world_size"] == 2 and md["failed_ranks"] == []
    assert sorted(md["passthrough"]) == ["norm.weight", "position_ids"]
    assert any(f.startswith("rank0_") for f in md["files"]) and any(f.startswith("rank1_") for f in md["files"])
    from safetensors.torch import load_file
    flat = {}
    for f in md["files"]:
        flat.update(load_file(os.path.join(out, f)))
    for n in expect:
        want = O.pack_result(O.group_quant_vec(tensors[n], 4, 128, False, False))
        assert torch.equal(flat[n + ".qweight"], want["qweight"]) and torch.equal(flat[n + ".qzeros"], want["qzeros"]), n
        assert torch.equal(flat[n + ".scales"].view(torch.int16), want["scales"].view(torch.int16)), n

#!/usr/bin/env python
"""The `awq_quantizer` CLI end to end on a synthetic checkpoint directory of a BASELINE.json model shape:
safetensors files on disk -> load_tensors() -> quantize (GPU) -> chunk files on disk.  Prints one JSON line
with the wall-clock split (the reference's own CLI on this input spends ~84 us per group in its Python loop:
SURVEY.md section 6).

    python tools/run_cli.py --workload opt-350m [--pack] [--save_safetensors]"""
import argparse, json, os, shutil, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "awq-converter_b200"))
import torch
from safetensors.torch import save_file
from awq_quantizer import model_shapes as M
from awq_quantizer import main as cli

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="opt-350m")
ap.add_argument("--pack", action="store_true")
ap.add_argument("--save_safetensors", action="store_true")
ap.add_argument("--packed_only", action="store_true")
ap.add_argument("--keep", action="store_true")
ap.add_argument("--search", action="store_true", help="write a calibration file (T tokens per searched linear) and pass --calibration_file")
ap.add_argument("--tokens", type=int, default=2048)
ap.add_argument("--runs", type=int, default=2)
args = ap.parse_args()

work = tempfile.mkdtemp(prefix="awq_cli_")
src, dst = os.path.join(work, "model"), os.path.join(work, "out")
os.makedirs(src)
gdev = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")      # synthetic data is generated where it is fast
gen = torch.Generator(device=gdev).manual_seed(7)
specs = M.workload(args.workload)
shard, nbytes, fi, total = {}, 0, 0, 0
t_gen = time.perf_counter()
for name, shape, _ in specs:
    shard[name] = (torch.randn(shape, generator=gen, device=gdev, dtype=torch.float32) * 0.02).to(torch.bfloat16).cpu()
    nbytes += shard[name].numel() * 2
    if nbytes > (2 << 30):
        save_file(shard, os.path.join(src, f"model-{fi:05d}.safetensors")); fi += 1; total += nbytes; shard, nbytes = {}, 0
if shard:
    save_file(shard, os.path.join(src, f"model-{fi:05d}.safetensors")); total += nbytes
del shard
argv = ["--model_id", src, "--output_dir", dst, "--log_level", "ERROR", "--chunk_size", "64"]
calib_bytes = 0
if args.search:                                   # {weight name: activations [T, K]} -- the CLI's calibration format
    acts, cache, alias = {}, {}, {}
    for name, shape, ck in specs:
        if ck is None or len(shape) != 2:
            continue
        key = (ck, shape[1])
        if key not in cache:                            # the first weight that sees this input stores it ...
            gain = torch.exp(torch.randn(shape[1], generator=gen, device=gdev))
            acts[name] = (torch.randn((args.tokens, shape[1]), generator=gen, device=gdev) * gain).to(torch.bfloat16).cpu()
            cache[key] = name
        else:                                           # ... the others refer to it (q/k/v, gate/up)
            alias["alias." + name] = cache[key]
    calib_path = os.path.join(work, "calibration.safetensors")
    save_file(acts, calib_path, metadata=alias)
    calib_bytes = os.path.getsize(calib_path)
    del acts, cache
    argv += ["--calibration_file", calib_path]
t_gen = time.perf_counter() - t_gen
if args.pack:
    argv.append("--pack")
if args.save_safetensors:
    argv.append("--save_safetensors")
if args.packed_only:
    argv.append("--packed_only")
runs = []
for it in range(args.runs):
    shutil.rmtree(dst, ignore_errors=True)
    t0 = time.perf_counter()
    rc = cli.main(argv)
    runs.append(time.perf_counter() - t0)
    assert rc == 0, rc
meta = json.load(open(os.path.join(dst, "metadata.json")))
out_bytes = sum(os.path.getsize(os.path.join(dst, f)) for f in os.listdir(dst))
print(json.dumps({"workload": args.workload, "cli": "awq_quantizer " + " ".join(argv[4:]), "bf16_GB": total / 1e9,
                  "calibration_GB": calib_bytes / 1e9, "synthetic_checkpoint_written_in_s": round(t_gen, 1),
                  "tensors_quantized": meta["num_tensors"], "wall_s_first_run": round(runs[0], 3),
                  "wall_s_second_run": round(runs[-1], 3) if len(runs) > 1 else None, "wall_s_all_runs": [round(r, 3) for r in runs],
                  "output_GB": out_bytes / 1e9, "timing_s_last_run": meta.get("timing_s_rank0"), "note": "load_tensors + quantize + save chunks, one process, 1 GPU"}))
if not args.keep:
    shutil.rmtree(work, ignore_errors=True)

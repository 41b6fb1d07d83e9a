"""``awq_quantizer`` CLI -- same flags, exit codes and output layout as the reference's main.py, re-hosted
on the B200 path:

* every flag of main.py:32-157 is accepted unchanged; additive flags: ``--pack`` (int32-packed
  qweight/qzeros next to the unpacked codes; ``--packed_only``: instead of them), ``--arith {native,fp32}``, ``--config``
  (the YAML file the reference documents but never wired, main.py:16 / USAGE.md:13-22),
  ``--calibration_file`` / ``--n_grid`` (activation-aware alpha search for the weights it names);
* tensor selection = main.py:243-253 (non-float / empty / numel < 128 are skipped, largest first);
* ``--multi_gpu`` no longer repeats the whole model on every device (main.py:596-606): tensors are
  partitioned largest-first onto the least-loaded device -- the reference's own, never-called,
  partition_tensors (main.py:395-427) -- and under ``torchrun`` each rank quantizes only its shard and
  writes its own chunk files; rank 0 gathers the chunk maps (NCCL/gloo all_gather_object) into
  metadata.json;
* ``--save_safetensors`` works (the reference hands nested dicts to save_file, main.py:478-490, and
  fails): flat keys ``{name}.q / .scales / .zero_points / .bits / .group_size / .symmetric`` as in
  test_quantization.py:182-189 (+ ``.qweight`` / ``.qzeros`` with --pack);
* no CPU execution: ``--device cpu`` (or no visible GPU) ends with exit code 1 and a clear message.
"""
from __future__ import annotations

import argparse
import json
import os
import time
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Optional, Tuple

import torch

from . import parallel
from .model_loading import load_model_from_hub
from .model_shapes import partition_lpt
from .quantization.awq import AWQQuantizer
from .utils.config import load_config
from .utils.logger import get_logger


def parse_args(argv=None) -> argparse.Namespace:
    p = argparse.ArgumentParser(description="AWQ Quantizer CLI")
    p.add_argument("--model_id", type=str, required=True, help="Model ID on HuggingFace Hub or path to local model")
    p.add_argument("--output_dir", type=str, required=True, help="Directory to save quantized model")
    p.add_argument("--bits", type=int, default=4, choices=[4, 8], help="Number of bits for quantization")
    p.add_argument("--group_size", type=int, default=128, help="Group size for quantization")
    p.add_argument("--symmetric", action="store_true", help="Use symmetric quantization")
    p.add_argument("--zero_point", type=str, default="minmax", choices=["none", "minmax", "percentile"],
                   help="Zero point calibration method")
    p.add_argument("--percentile", type=float, default=0.99, help="Percentile for zero point calibration")
    p.add_argument("--scale_method", type=str, default="mse", choices=["minmax", "mse"], help="Scale calibration method")
    p.add_argument("--per_channel", action="store_true", help="Use per-channel quantization")
    p.add_argument("--device", type=str, default="cuda" if torch.cuda.is_available() else "cpu",
                   help="Device to use for quantization (cuda, cuda:0, cuda:1, cpu, or 'all' for all GPUs)")
    p.add_argument("--num_workers", type=int, default=4, help="Number of worker threads for parallel processing per GPU")
    p.add_argument("--max_memory", type=float, default=0.8, help="Maximum fraction of GPU memory to use (0.0-1.0)")
    p.add_argument("--multi_gpu", action="store_true", help="Use all available GPUs for processing (overrides --device)")
    p.add_argument("--batch_size", type=int, default=10, help="Number of tensors to process in each batch")
    p.add_argument("--prefetch_factor", type=int, default=2, help="Number of batches to prefetch")
    p.add_argument("--memory_efficient", action="store_true", help="Enable memory-efficient mode")
    p.add_argument("--log_level", type=str, default="INFO", choices=["DEBUG", "INFO", "WARNING", "ERROR", "CRITICAL"],
                   help="Logging level")
    p.add_argument("--log_file", type=str, help="Log file path")
    p.add_argument("--save_safetensors", action="store_true", help="Save in safetensors format instead of pytorch format")
    p.add_argument("--chunk_size", type=int, default=10, help="Number of tensors to save in each chunk (for large models)")
    # additive
    p.add_argument("--pack", action="store_true", help="emit int32-packed qweight/qzeros (8 nibbles per word)")
    p.add_argument("--packed_only", action="store_true",
                   help="with --pack: leave out the unpacked int32 codes (the reference's `tensor_q` / `.q`, 4 bytes "
                        "per weight: 8x the packed size on the PCIe link and on disk)")
    p.add_argument("--arith", type=str, default="native", choices=["native", "fp32"],
                   help="arithmetic contract: native = the reference's own (input dtype), fp32 = reference on w.float()")
    p.add_argument("--config", type=str, help="YAML config (reference schema); CLI flags win over it")
    p.add_argument("--calibration_file", type=str,
                   help="safetensors file {weight name: activations [tokens, in_features]}: run the "
                        "activation-aware alpha search (scale_method) for those weights; __metadata__ entries "
                        "'alias.<weight name>' = '<stored name>' let several weights share one stored tensor")
    p.add_argument("--n_grid", type=int, default=20, help="points of the alpha grid for --calibration_file")
    return p.parse_args(argv)


def get_available_gpus(logger=None) -> List[str]:
    if not torch.cuda.is_available():
        if logger:
            logger.warning("No CUDA devices available")
        return []
    out = []
    for i in range(torch.cuda.device_count()):
        if logger:
            logger.info(f"Found CUDA device {i}: {torch.cuda.get_device_name(i)}")
        out.append(f"cuda:{i}")
    return out


def get_device_memory_info(device_idx: int) -> Tuple[float, float]:
    if not torch.cuda.is_available():
        return 0.0, 0.0
    try:
        free, total = torch.cuda.mem_get_info(device_idx)
        return total / 1024 ** 3, free / 1024 ** 3
    except Exception:
        return 0.0, 0.0


def prepare_tensors_for_quantization(tensors: Dict[str, torch.Tensor], device: str, max_memory_fraction: float = 0.8,
                                     batch_size: int = 10, logger=None) -> List[Dict[str, torch.Tensor]]:
    """main.py:216-330: skip rules, largest-first order, batches of <= batch_size tensors that fit the
    GPU-memory budget.  Tensors stay on the host; the quantizer uploads them (pipelined)."""
    items = []
    for name, t in tensors.items():
        if not isinstance(t, torch.Tensor) or not t.is_floating_point() or t.numel() == 0:
            if logger:
                logger.warning(f"Skipping invalid tensor: {name}")
            continue
        if t.numel() < 128:
            if logger:
                logger.warning(f"Skipping tensor too small for grouping: {name}")
            continue
        items.append((name, t, t.numel() * t.element_size()))
    items.sort(key=lambda x: x[2], reverse=True)
    budget = None
    if device.startswith("cuda") and torch.cuda.is_available():
        idx = int(device.split(":")[1]) if ":" in device else torch.cuda.current_device()
        total = torch.cuda.get_device_properties(idx).total_memory
        budget = int(total * max_memory_fraction) - torch.cuda.memory_allocated(idx)
    batches, cur, cur_bytes = [], {}, 0
    for name, t, nbytes in items:
        # x3: input + int32 codes + slack -- a batch must fit on the device at once
        if cur and (len(cur) >= batch_size or (budget is not None and cur_bytes + 3 * nbytes > budget)):
            batches.append(cur)
            cur, cur_bytes = {}, 0
        cur[name] = t
        cur_bytes += 3 * nbytes
    if cur:
        batches.append(cur)
    if logger:
        logger.info(f"Created {len(batches)} batches with {sum(len(b) for b in batches)} total tensors")
    return batches


def quantize_tensor_batch(tensor_items: List[tuple], quantizer: AWQQuantizer, device: str, logger=None,
                          pack: bool = False) -> Dict[str, dict]:
    """main.py:333-392 (without its OOM -> CPU fallback: failures are logged and the tensor is skipped)."""
    out = {}
    for name, tensor in tensor_items:
        if logger:
            logger.info(f"Quantizing tensor: {name} on {device}")
        try:
            out[name] = quantizer.quantize(tensor, pack=pack)
            if logger:
                logger.info(f"Successfully quantized tensor: {name} on {device}")
        except Exception as e:
            if logger:
                logger.error(f"Failed to quantize tensor {name} on {device}: {e}")
    return out


def partition_tensors(tensors: Dict[str, torch.Tensor], num_partitions: int) -> List[Dict[str, torch.Tensor]]:
    """main.py:395-427 -- largest first onto the least-loaded partition, by bytes."""
    if num_partitions <= 1:
        return [tensors]
    bins = partition_lpt([(n, t.numel() * t.element_size()) for n, t in tensors.items()], num_partitions)
    return [{n: tensors[n] for n in b} for b in bins]


def _own_storage(qd: dict) -> dict:
    """torch.save pickles the WHOLE storage behind a view; the pipeline returns views of model-sized result
    arrays, so every tensor is detached into its own storage first (one host copy of the results)"""
    out = {}
    for k, v in qd.items():
        if isinstance(v, torch.Tensor) and v.numel() * v.element_size() != v.untyped_storage().nbytes():
            v = v.clone()
        out[k] = v
    return out


def _flatten_for_safetensors(chunk: Dict[str, dict]) -> Dict[str, torch.Tensor]:
    rename = {"tensor_q": "q"}
    flat = {}
    for name, qd in chunk.items():
        for k, v in qd.items():
            if isinstance(v, torch.Tensor):
                flat[f"{name}.{rename.get(k, k)}"] = v.contiguous() if v.dim() else v.reshape(1)
    return flat


_ST_DTYPES = {torch.float16: "F16", torch.bfloat16: "BF16", torch.float32: "F32", torch.float64: "F64",
              torch.int64: "I64", torch.int32: "I32", torch.int16: "I16", torch.int8: "I8", torch.uint8: "U8",
              torch.bool: "BOOL"}


def write_safetensors(flat: Dict[str, torch.Tensor], path: str) -> None:
    """One safetensors file written STRAIGHT from the tensors' own host memory (the result arenas of the
    pipelines): 8-byte header length, JSON header padded to 8 bytes, then the raw little-endian data in header
    order.  No intermediate serialisation buffer, and the write() calls release the GIL, so the chunk files of a
    model are written by several threads at once (safetensors.torch.save_file does neither)."""
    header, off, bufs = {}, 0, []
    for name in sorted(flat):                                    # (any order is valid: offsets must only be gap-free)
        t = flat[name]
        if t.dtype not in _ST_DTYPES:
            raise ValueError(f"{name}: dtype {t.dtype} has no safetensors code")
        t = t.detach()
        if t.device.type != "cpu":
            t = t.cpu()
        t = t.contiguous()
        nb = t.numel() * t.element_size()
        header[name] = {"dtype": _ST_DTYPES[t.dtype], "shape": list(t.shape), "data_offsets": [off, off + nb]}
        off += nb
        if nb:
            bufs.append(t.reshape(-1).view(torch.uint8).numpy())
    blob = json.dumps(header, separators=(",", ":")).encode()
    blob += b" " * (-len(blob) % 8)
    with open(path, "wb", buffering=0) as f:
        f.write(len(blob).to_bytes(8, "little"))
        f.write(blob)
        for b in bufs:
            mv = memoryview(b)
            done = 0
            while done < len(mv):                                # (a single write() moves at most 2 GiB on Linux)
                done += f.write(mv[done:done + (1 << 30)])


def save_model_in_chunks(tensors: Dict[str, dict], output_dir: str, chunk_size: int = 10, use_safetensors: bool = False,
                         logger=None, rank: int = 0, world: int = 1, write_metadata: bool = True) -> dict:
    """main.py:430-512.  Chunk files ``model_chunk_%04d`` (+ ``rankN_`` prefix when sharded over ranks)."""
    os.makedirs(output_dir, exist_ok=True)
    names = list(tensors.keys())
    num_chunks = (len(names) + chunk_size - 1) // chunk_size
    prefix = f"rank{rank}_" if world > 1 else ""
    example = tensors[names[0]] if names else {}
    params = {k: (example[k].item() if k in example else None) for k in ("bits", "group_size", "symmetric")}
    tensor_to_chunk, files = {}, []

    def write_chunk(c: int) -> str:
        part = {n: tensors[n] for n in names[c * chunk_size:(c + 1) * chunk_size]}
        base = os.path.join(output_dir, f"{prefix}model_chunk_{c:04d}")
        if use_safetensors:
            write_safetensors(_flatten_for_safetensors(part), base + ".safetensors")
            out = os.path.basename(base) + ".safetensors"
        else:
            torch.save({n: _own_storage(qd) for n, qd in part.items()}, base + ".pt")
            out = os.path.basename(base) + ".pt"
        if logger:
            logger.info(f"Saved chunk {c + 1}/{num_chunks} with {len(part)} tensors")
        return out

    for c in range(num_chunks):
        for n in names[c * chunk_size:(c + 1) * chunk_size]:
            tensor_to_chunk[n] = c
    # chunk files are independent: several writer threads (the file writes release the GIL; one thread reaches
    # ~2 GB/s into the page cache, a 70B-shaped shard is 4.6 GB of packed output per rank)
    workers = min(num_chunks, max(1, min(8, (os.cpu_count() or 4) // max(1, world))))
    if workers > 1:
        with ThreadPoolExecutor(max_workers=workers) as ex:
            files = list(ex.map(write_chunk, range(num_chunks)))
    else:
        files = [write_chunk(c) for c in range(num_chunks)]
    meta = {"rank": rank, "num_chunks": num_chunks, "chunk_size": chunk_size, "tensor_to_chunk": tensor_to_chunk,
            "format": "safetensors" if use_safetensors else "pytorch", "num_tensors": len(names),
            "quantization_params": params, "files": files}
    if write_metadata and world == 1:
        with open(os.path.join(output_dir, "metadata.json"), "w") as f:
            json.dump({k: v for k, v in meta.items() if k != "rank"}, f, indent=2)
    return meta


def _apply_config(args: argparse.Namespace, argv) -> None:
    if not args.config:
        return
    cfg = load_config(args.config)
    given = {a.split("=")[0].lstrip("-") for a in (argv or []) if a.startswith("--")}
    for key in ("bits", "group_size", "symmetric", "zero_point", "percentile", "scale_method", "per_channel"):
        if key not in given and cfg.get(f"quantization.{key}") is not None:
            setattr(args, key, cfg.get(f"quantization.{key}"))
    args.skip_layers = list(cfg.get("quantization.skip_layers", []) or [])


_ALIAS = "alias."


def read_calibration_index(path: str) -> Dict[str, int]:
    """header only: {weight name: calibration tokens}.  A calibration file maps weight names to activation tensors
    [tokens, in_features]; linears that see the same input (q/k/v, gate/up) may share ONE stored tensor through
    ``__metadata__`` entries ``alias.<weight name> = <stored tensor name>``."""
    from safetensors import safe_open
    with safe_open(path, framework="pt") as f:
        tokens = {k: f.get_slice(k).get_shape()[0] for k in f.keys()}
        for k, v in (f.metadata() or {}).items():
            if k.startswith(_ALIAS) and v in tokens:
                tokens[k[len(_ALIAS):]] = tokens[v]
    return tokens


def load_calibration(path: str) -> Dict[str, torch.Tensor]:
    """{weight name: activations}; aliased weights get the SAME tensor object, so that the streamed search uploads
    it once and computes its scale grid once"""
    from safetensors import safe_open
    from safetensors.torch import load_file
    calib = load_file(path)
    with safe_open(path, framework="pt") as f:
        for k, v in (f.metadata() or {}).items():
            if k.startswith(_ALIAS) and v in calib:
                calib[k[len(_ALIAS):]] = calib[v]
    return calib


def select_tensors(index, skip_layers, logger=None):
    """main.py:243-253 on the header index: (quantizable names largest first, pass-through names).  Non-float,
    empty and numel < 128 tensors -- and those matching ``quantization.skip_layers`` (default_config.yaml:35) -- are
    not quantized; the reference drops them, here they are passed through (bf16 -> fp16 by K3, others unchanged)."""
    quant, passthrough = [], []
    for name, info in index.items():
        floating = info.dtype is not None and info.dtype.is_floating_point
        if any(s in name for s in skip_layers):
            passthrough.append(name)
        elif not floating or info.numel == 0:
            if logger:
                logger.warning(f"Skipping invalid tensor: {name}")
            passthrough.append(name)
        elif info.numel < 128:
            if logger:
                logger.warning(f"Skipping tensor too small for grouping: {name}")
            passthrough.append(name)
        else:
            quant.append(name)
    quant.sort(key=lambda n: index[n].nbytes, reverse=True)
    return quant, passthrough


def convert_passthrough(tensors: Dict[str, torch.Tensor], device: str) -> Dict[str, torch.Tensor]:
    """the tensors that are NOT quantized: bf16 -> fp16 through K3 (tensor_utils.py:10-22, the conversion the
    reference's loader offers at safetensors_loader.py:205-225), everything else unchanged"""
    from .utils.tensor_utils import convert_bf16_to_fp16
    return {n: (convert_bf16_to_fp16(t, device=device) if t.dtype == torch.bfloat16 and t.numel() else t)
            for n, t in tensors.items()}


def save_passthrough(tensors: Dict[str, torch.Tensor], output_dir: str, rank: int, world: int) -> Optional[str]:
    if not tensors:
        return None
    from safetensors.torch import save_file
    fn = (f"rank{rank}_" if world > 1 else "") + "passthrough.safetensors"
    save_file({n: t.contiguous() for n, t in tensors.items()}, os.path.join(output_dir, fn))
    return fn


def _rank_work(args, logger, rank: int, world: int, local: int) -> dict:
    """Everything one process does; returns its metadata record (with 'error' set on failure) and never raises:
    under torchrun every rank must reach the metadata gather, or the others would wait in it."""
    meta = {"rank": rank, "num_chunks": 0, "chunk_size": args.chunk_size, "tensor_to_chunk": {}, "num_tensors": 0,
            "format": "safetensors" if args.save_safetensors else "pytorch", "quantization_params": None, "files": [],
            "passthrough": {}, "error": None}
    try:
        if world > 1:
            devices = [f"cuda:{parallel.local_device_index(local)}"]
        elif args.multi_gpu or args.device.lower() == "all":
            devices = get_available_gpus(logger)
        else:
            devices = [args.device]
        if not devices or any(not d.startswith("cuda") for d in devices) or not torch.cuda.is_available():
            meta["error"] = ("This build of awq_quantizer runs on CUDA (B200) only; no CPU execution path exists "
                             f"(requested devices: {devices or 'none found'})")
            return meta
        for d in devices:
            idx = int(d.split(":")[1]) if ":" in d else torch.cuda.current_device()
            tot, free = get_device_memory_info(idx)
            logger.info(f"Using GPU {d}: {torch.cuda.get_device_name(idx)} ({tot:.1f} GB total, {free:.1f} GB free)")

        logger.info(f"Loading model from {args.model_id}")
        t_phase = time.perf_counter()
        timing = meta.setdefault("timing_s", {})

        def lap(what: str) -> None:
            nonlocal t_phase
            now = time.perf_counter()
            timing[what] = round(timing.get(what, 0.0) + now - t_phase, 3)
            t_phase = now

        try:
            loader = load_model_from_hub(args.model_id, logger_level=args.log_level)
            index = loader.index()                       # headers only: nothing is read before the partition is known
        except Exception as e:
            meta["error"] = f"Failed to load model: {e}"
            return meta
        quant_names, pass_names = select_tensors(index, args.skip_layers, logger)
        calib_tokens = {}
        if args.calibration_file and args.scale_method == "mse":
            try:                                         # header only: which weights will be searched, and over how many tokens
                calib_tokens = read_calibration_index(args.calibration_file)
            except Exception as e:
                meta["error"] = f"Failed to read calibration file: {e}"
                return meta
        if world > 1:                                    # rank-local shard (torchrun), largest first: cost = bytes
            # (main.py:410), or C*K*T for the linears that run the activation-aware search (tensor-core bound)
            def cost(n):
                info = index[n]
                if n in calib_tokens and len(info.shape) == 2:
                    return info.numel * max(2, int(calib_tokens[n]))
                return info.nbytes
            costs = [(n, cost(n)) for n in quant_names] + [(n, index[n].nbytes) for n in pass_names]
            mine = set(parallel.shard_for_rank(costs, world, rank))
            quant_names = [n for n in quant_names if n in mine]
            pass_names = [n for n in pass_names if n in mine]
        try:                                             # only this rank's tensors, as views of the mapped files
            tensors = loader.load_tensors(names=quant_names + pass_names)
        except Exception as e:
            meta["error"] = f"Failed to load model: {e}"
            return meta
        quantizable = {n: tensors[n] for n in quant_names}
        shards = [quantizable] if world > 1 else partition_tensors(quantizable, len(devices))
        lap("index_and_map")

        calib = {}
        if args.calibration_file:
            if args.scale_method != "mse":
                logger.warning("--calibration_file is ignored: the activation-aware search belongs to scale_method=mse")
            else:
                calib = load_calibration(args.calibration_file)
                logger.info(f"Loaded calibration activations for {len(calib)} tensors")
        lap("map_calibration")

        def run_device(device: str, shard: Dict[str, torch.Tensor]) -> Dict[str, dict]:
            qz = AWQQuantizer(bits=args.bits, group_size=args.group_size, symmetric=args.symmetric,
                              zero_point=args.zero_point, percentile=args.percentile, scale_method=args.scale_method,
                              per_channel=args.per_channel, device=device, logger_name=f"awq_quantizer_{device}",
                              logger_level=args.log_level, logger_to_file=args.log_file is not None,
                              logger_file_path=args.log_file, arith=args.arith, n_grid=args.n_grid)
            for name in shard:
                logger.info(f"Quantizing tensor: {name} on {device}")
            # the whole shard in ONE call: searched linears stream in waves (quantization/stream.py), every other
            # whole-group tensor through the native gather pipeline (bounded pinned rings, no per-tensor upload /
            # kernel / download round trip as in main.py:333-392); what neither can take (ragged rows ...) falls
            # back to per-tensor calls inside.  --batch_size / --num_workers / --prefetch_factor are accepted for
            # compatibility and not needed.
            acts = {n: calib[n] for n, t in shard.items() if n in calib and t.dim() == 2}
            try:
                return qz.quantize_model(shard, activations=acts or None, pack=args.pack,
                                         keep_unpacked=not (args.pack and args.packed_only))
            except Exception as e:
                logger.error(f"Error processing shard on {device}: {e}")
                return {}

        quantized: Dict[str, dict] = {}
        if len(devices) == 1:
            quantized = run_device(devices[0], shards[0])
        else:
            with ThreadPoolExecutor(max_workers=len(devices)) as ex:
                for part in ex.map(lambda ds: run_device(*ds), zip(devices, shards)):
                    quantized.update(part)
        logger.info(f"Successfully quantized {len(quantized)} tensors")
        lap("quantize")
        try:
            saved = save_model_in_chunks(quantized, args.output_dir, args.chunk_size, args.save_safetensors, logger,
                                         rank=rank, world=world, write_metadata=False)
            passed = convert_passthrough({n: tensors[n] for n in pass_names}, devices[0])
            fn = save_passthrough(passed, args.output_dir, rank, world)
        except Exception as e:
            meta["error"] = f"Failed to save quantized model: {e}"
            return meta
        lap("save")
        meta.update(saved)
        meta["passthrough"] = {n: fn for n in passed}
        return meta
    except Exception as e:                               # anything unexpected still reaches the gather
        meta["error"] = f"Error during quantization: {e}"
        return meta


def main(argv=None) -> int:
    logger = None
    dist_up = False
    try:
        import sys
        argv = sys.argv[1:] if argv is None else list(argv)
        args = parse_args(argv)
        args.skip_layers = []
        _apply_config(args, argv)
        logger = get_logger(name="awq_quantizer", level=args.log_level, to_file=args.log_file is not None,
                            file_path=args.log_file)
        os.makedirs(args.output_dir, exist_ok=True)
        rank, world = parallel.init_distributed()
        dist_up = world > 1
        local = parallel.rank_info()[2]
        if world > 1 and torch.cuda.is_available():
            parallel.bind_to_gpu_numa(parallel.local_device_index(local))
        start = time.time()
        meta = _rank_work(args, logger, rank, world, local)
        if meta["error"]:
            logger.error(meta["error"])
        metas = parallel.gather_metadata(meta)           # every rank arrives here, failed or not
        errors = [m for m in metas if m.get("error")]
        merged = parallel.merge_chunk_maps([m for m in metas if not m.get("error")])
        if rank == 0 and not (world == 1 and errors):
            passthrough = {}
            for m in metas:
                passthrough.update(m.get("passthrough") or {})
            merged.update({"chunk_size": args.chunk_size, "format": meta["format"], "timing_s_rank0": meta.get("timing_s"),
                           "quantization_params": next((m["quantization_params"] for m in metas
                                                        if m.get("num_tensors") and m.get("quantization_params")), None),
                           "passthrough": passthrough})
            if world > 1:
                merged["world_size"] = world
                merged["failed_ranks"] = [m["rank"] for m in errors]
            with open(os.path.join(args.output_dir, "metadata.json"), "w") as f:
                json.dump(merged, f, indent=2)
        if errors:
            return 1
        if not merged["num_tensors"]:
            logger.error("No tensors were successfully quantized")
            return 1
        logger.info(f"Quantization complete in {time.time() - start:.2f} seconds")
        return 0
    except SystemExit:
        raise
    except Exception as e:
        if logger is not None:
            logger.error(f"Error during quantization: {e}")
        else:
            print(f"Error during quantization: {e}")
        return 1
    finally:
        if dist_up:
            try:
                import torch.distributed as dist
                if dist.is_initialized():
                    dist.destroy_process_group()
            except Exception:
                pass


if __name__ == "__main__":
    raise SystemExit(main())

"""Rank-local multi-GPU plumbing.  The hot path shards by tensor with no data-path collective
(every group of every row is independent, awq.py:332-368); the only communication is a final gather
of per-rank metadata so that rank 0 can write metadata.json (main.py:499-509).

One process per GPU (torchrun): NCCL on GPUs, gloo on CPU (tests).  The partitioner is the
reference's own -- never called -- partition_tensors (main.py:395-427)."""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

from .model_shapes import partition_lpt


def rank_info() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1-process defaults)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_distributed(backend: Optional[str] = None):
    """Initialises torch.distributed when WORLD_SIZE > 1 (idempotent).  Returns (rank, world)."""
    import torch
    import torch.distributed as dist
    rank, world, local = rank_info()
    if world <= 1:
        return 0, 1
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            # AWQ_DIST_BACKEND=gloo: ranks that share a GPU (fewer devices than ranks; NCCL refuses duplicates)
            backend = os.environ.get("AWQ_DIST_BACKEND") or ("nccl" if torch.cuda.is_available() else "gloo")
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, rank=rank, world_size=world, **kwargs)
    return rank, world


def local_device_index(local_rank: int) -> int:
    """the CUDA device of a local rank: its own GPU; ranks beyond the device count wrap around (sharing a GPU
    works for this path -- no collective runs on the device -- and is what a 1-GPU test box needs)"""
    import torch
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    return local_rank % n if n else local_rank


def bind_to_gpu_numa(device_index: int) -> Optional[int]:
    """Best effort: pin this process to the CPUs of the NUMA node the GPU hangs off, BEFORE pinned host
    buffers are allocated (first-touch places them on that node; H2D/D2H then stay off the inter-socket
    link).  Matters when 8 ranks stream their arenas at once.  Returns the node or None."""
    try:
        import torch
        props = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, props.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


def shard_for_rank(items: Sequence[Tuple[str, int]], world: int, rank: int) -> List[str]:
    """names owned by `rank` under the deterministic LPT partition of (name, cost) items"""
    return partition_lpt(list(items), world)[rank]


def tensor_costs(tensors: Dict[str, "object"], searched_tokens: int = 0) -> List[Tuple[str, int]]:
    """cost model: bytes (main.py:410), or C*K*T for linears that run the activation-aware search"""
    out = []
    for name, t in tensors.items():
        n = t.numel()
        cost = n * t.element_size()
        if searched_tokens and t.dim() == 2:
            cost = n * searched_tokens
        out.append((name, cost))
    return out


def gather_metadata(local: dict) -> List[dict]:
    """all ranks -> every rank: one all_gather_object of small dicts (KBs)"""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [local]
    out: List[Optional[dict]] = [None] * dist.get_world_size()
    dist.all_gather_object(out, local)
    return out  # type: ignore[return-value]


def merge_chunk_maps(per_rank: List[dict]) -> dict:
    """per-rank {'rank', 'num_chunks', 'tensor_to_chunk'} -> global metadata (chunk ids offset by rank order)"""
    tensor_to_chunk, offset, files = {}, 0, []
    for meta in sorted(per_rank, key=lambda m: m["rank"]):
        for name, c in meta["tensor_to_chunk"].items():
            if name in tensor_to_chunk:
                raise ValueError(f"tensor {name} quantized by two ranks")
            tensor_to_chunk[name] = offset + c
        offset += meta["num_chunks"]
        files.extend(meta.get("files", []))
    return {"num_chunks": offset, "tensor_to_chunk": tensor_to_chunk, "num_tensors": len(tensor_to_chunk),
            "files": files}

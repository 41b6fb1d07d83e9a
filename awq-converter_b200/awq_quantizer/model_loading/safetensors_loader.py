"""SafetensorsLoader -- same constructor and methods as the reference
(model_loading/safetensors_loader.py:17-225).  ``load_tensors()`` keeps its contract (Dict[str, CPU
tensor], stored dtype, later shards overwrite duplicates with a warning).  Added for the B200 path
(SURVEY.md section 8f row 2):

* ``index()``                  name -> TensorInfo (file, shape, dtype, bytes) from the file HEADERS only -- what the
                               orchestrator partitions over the ranks before anything is read;
* ``load_tensors(names=...)``  only the named tensors, as views of the memory-mapped files (``safe_open``): a
                               rank touches only the pages of its own shard, and the pages are read from disk
                               while the gather pipeline copies them into its pinned ring -- file read, H2D and
                               the kernels overlap.  (``load_tensors()`` without names keeps the reference's
                               eager whole-file ``load_file``.)
* ``load_arena()``             every arena-eligible tensor straight into ONE pinned, tile-aligned host arena
                               (quantization/arena.py)."""
from __future__ import annotations

import os
from typing import Dict, Iterable, NamedTuple, Optional, Tuple

import torch

from ..utils.logger import get_logger
from ..utils.tensor_utils import convert_bf16_to_fp16, filter_consolidated_files, get_model_files


_ST_DTYPES = {"BF16": torch.bfloat16, "F16": torch.float16, "F32": torch.float32, "F64": torch.float64,
              "I64": torch.int64, "I32": torch.int32, "I16": torch.int16, "I8": torch.int8, "U8": torch.uint8,
              "BOOL": torch.bool}


class TensorInfo(NamedTuple):
    path: str
    shape: Tuple[int, ...]
    dtype: Optional[torch.dtype]       # None: a storage type torch has no name for here (passed through by load_file)
    nbytes: int

    @property
    def numel(self) -> int:
        n = 1
        for s in self.shape:
            n *= s
        return n


class SafetensorsLoader:
    def __init__(self, model_path: str, from_hub: bool = False, revision: str = "main", token: Optional[str] = None,
                 logger_name: str = "safetensors_loader", logger_level: str = "INFO", logger_to_file: bool = False,
                 logger_file_path: Optional[str] = None, resume_download: bool = True, force_download: bool = False):
        self.model_path = model_path
        self.from_hub = from_hub
        self.revision = revision
        self.token = token
        self.resume_download = resume_download
        self.force_download = force_download
        self.logger = get_logger(name=logger_name, level=logger_level, to_file=logger_to_file,
                                 file_path=logger_file_path)
        self.model_files = get_model_files(model_path)
        if not self.model_files:
            raise ValueError(f"No safetensor files found in {model_path}")
        self.model_files = filter_consolidated_files(self.model_files)
        self.logger.info(f"Loading {len(self.model_files)} safetensors files:")
        for f in self.model_files:
            self.logger.info(f"  - {os.path.basename(f)}")
        self.tensors: Dict[str, torch.Tensor] = {}

    def verify_file(self, file_path: str) -> bool:
        try:
            from safetensors import safe_open
            with safe_open(file_path, framework="pt") as f:
                return len(list(f.keys())) >= 0
        except Exception as e:
            self.logger.error(f"File verification failed for {file_path}: {e}")
            return False

    def index(self) -> Dict[str, TensorInfo]:
        """headers only: no tensor data is read.  Later files win on duplicate names, like load_tensors()."""
        from safetensors import safe_open
        out: Dict[str, TensorInfo] = {}
        for path in self.model_files:
            with safe_open(path, framework="pt") as f:
                for name in f.keys():
                    sl = f.get_slice(name)
                    shape = tuple(sl.get_shape())
                    dt = _ST_DTYPES.get(sl.get_dtype())
                    n = 1
                    for d in shape:
                        n *= d
                    if name in out:
                        self.logger.warning(f"Duplicate tensor name: {name}")
                    out[name] = TensorInfo(path, shape, dt, n * (torch.empty((), dtype=dt).element_size() if dt else 0))
        return out

    def _load_named(self, names: Iterable[str]) -> Dict[str, torch.Tensor]:
        from safetensors import safe_open
        want = set(names)
        idx = self.index()
        by_file: Dict[str, list] = {}
        for n in want:
            if n not in idx:
                raise KeyError(f"tensor {n} is in none of the safetensors files")
            by_file.setdefault(idx[n].path, []).append(n)
        got: Dict[str, torch.Tensor] = {}
        for path in self.model_files:                       # file order = the order load_tensors() reads in
            if path not in by_file:
                continue
            with safe_open(path, framework="pt") as f:       # tensors are views of the file mapping (lazy pages)
                for n in by_file[path]:
                    got[n] = f.get_tensor(n)
        self.tensors = {n: got[n] for n in idx if n in got}
        self.logger.info(f"Loaded {len(self.tensors)} of {len(idx)} tensors")
        return self.tensors

    def load_tensors(self, names: Optional[Iterable[str]] = None) -> Dict[str, torch.Tensor]:
        if names is not None:
            return self._load_named(names)
        from safetensors.torch import load_file
        self.tensors = {}
        for path in self.model_files:
            try:
                part = load_file(path)
            except Exception as e:
                self.logger.error(f"Error loading {os.path.basename(path)}: {e}")
                raise
            for name, t in part.items():
                if name in self.tensors:
                    self.logger.warning(f"Duplicate tensor name: {name}")
                self.tensors[name] = t
            self.logger.debug(f"Loaded {len(part)} tensors from {os.path.basename(path)}")
        self.logger.info(f"Loaded {len(self.tensors)} total tensors")
        return self.tensors

    def load_arena(self, group_size: int = 128, bits: int = 4, names=None):
        """(HostArena of the arena-eligible tensors, dict of the remaining tensors)."""
        from safetensors import safe_open
        from ..quantization.arena import HostArena, arena_eligible
        str2dt = {"BF16": torch.bfloat16, "F16": torch.float16, "F32": torch.float32}
        where, specs, rest = {}, {}, {}
        for path in self.model_files:
            with safe_open(path, framework="pt") as f:
                for name in f.keys():
                    if names is not None and name not in names:
                        continue
                    sl = f.get_slice(name)
                    shape, dt = tuple(sl.get_shape()), str2dt.get(sl.get_dtype())
                    if dt is not None and arena_eligible(shape, dt, group_size, bits):
                        specs[name] = (shape, dt)
                        where[name] = path
                    else:
                        rest[name] = f.get_tensor(name)
        arena = HostArena(specs) if specs else None
        by_file: Dict[str, list] = {}
        for name, path in where.items():
            by_file.setdefault(path, []).append(name)
        for path, ns in by_file.items():
            with safe_open(path, framework="pt") as f:
                for name in ns:
                    arena.views[name].copy_(f.get_tensor(name))
        return arena, rest

    def save_tensors(self, tensors: Dict[str, torch.Tensor], output_dir: str, filename: str = "model.safetensors") -> None:
        from safetensors.torch import save_file
        try:
            os.makedirs(output_dir, exist_ok=True)
            out = os.path.join(output_dir, filename)
            save_file({k: v.contiguous() for k, v in tensors.items()}, out)
            self.logger.info(f"Saved {len(tensors)} tensors to {out}")
        except Exception as e:
            self.logger.error(f"Error saving tensors: {e}")
            raise

    def convert_tensors_bf16_to_fp16(self, tensors: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        out = {}
        for name, t in tensors.items():
            c = convert_bf16_to_fp16(t)
            if c.dtype != t.dtype:
                self.logger.debug(f"Converted tensor {name} from {t.dtype} to {c.dtype}")
            out[name] = c
        return out

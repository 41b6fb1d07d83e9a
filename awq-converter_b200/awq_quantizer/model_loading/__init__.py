"""Model loading: the reference's surface (model_loading/__init__.py:18-151) over local safetensors
files.  I/O only -- no arithmetic lives here.  Hub access needs a network; a local directory passed as
the model id works offline exactly as it does in the reference (snapshot_download fails -> the path is
walked, model_loading/__init__.py:86-104)."""
from __future__ import annotations

import hashlib
import os
from typing import Optional

from ..utils.logger import get_logger
from .safetensors_loader import SafetensorsLoader

__all__ = ["load_model_from_hub", "load_model_from_path", "SafetensorsLoader", "verify_file_hash"]


def verify_file_hash(file_path: str, expected_hash: Optional[str] = None) -> str:
    h = hashlib.sha256()
    with open(file_path, "rb") as f:
        for block in iter(lambda: f.read(1 << 20), b""):
            h.update(block)
    digest = h.hexdigest()
    if expected_hash is not None and digest != expected_hash:
        raise ValueError(f"File hash mismatch for {file_path}. Expected: {expected_hash}, got: {digest}")
    return digest


def load_model_from_hub(model_id: str, revision: str = "main", token: Optional[str] = None,
                        logger_name: str = "model_loading", logger_level: str = "INFO",
                        logger_to_file: bool = False, logger_file_path: Optional[str] = None,
                        resume_download: bool = True, force_download: bool = False,
                        verify_downloads: bool = True) -> SafetensorsLoader:
    logger = get_logger(name=logger_name, level=logger_level, to_file=logger_to_file, file_path=logger_file_path)
    local_dir = None
    if not os.path.exists(model_id):
        try:
            from huggingface_hub import snapshot_download
            local_dir = snapshot_download(repo_id=model_id, revision=revision, token=token,
                                          force_download=force_download)
            logger.info(f"Successfully downloaded model snapshot to {local_dir}")
        except Exception as e:  # offline / invalid id: fall through to the path itself, as the reference does
            logger.warning(f"Failed to download model snapshot: {e}")
    return SafetensorsLoader(model_path=local_dir or model_id, from_hub=local_dir is None and not os.path.exists(model_id),
                             revision=revision, token=token, logger_name=logger_name, logger_level=logger_level,
                             logger_to_file=logger_to_file, logger_file_path=logger_file_path,
                             resume_download=resume_download, force_download=force_download)


def load_model_from_path(model_path: str, logger_name: str = "model_loading", logger_level: str = "INFO",
                         logger_to_file: bool = False, logger_file_path: Optional[str] = None,
                         verify_files: bool = True) -> SafetensorsLoader:
    return SafetensorsLoader(model_path=model_path, from_hub=False, logger_name=logger_name,
                             logger_level=logger_level, logger_to_file=logger_to_file,
                             logger_file_path=logger_file_path, resume_download=False, force_download=False)

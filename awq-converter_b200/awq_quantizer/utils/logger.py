"""Logging wrapper with the reference's surface (utils/logger.py:11-103): ``get_logger(name, level,
to_file, file_path)`` returns an object with debug/info/warning/error/critical; same record format;
handlers of a same-named logger are replaced on construction.  Unlike the reference it also exposes
``.level`` (main.py:313 reads it)."""
from __future__ import annotations

import logging
import os
import sys
from typing import Optional

_FORMAT = "%(asctime)s - %(name)s - %(levelname)s - %(message)s"


class Logger:
    def __init__(self, name: str = "awq_quantizer", level: str = "INFO", to_file: bool = False,
                 file_path: Optional[str] = None):
        lvl = getattr(logging, str(level).upper())
        self.logger = logging.getLogger(name)
        self.logger.setLevel(lvl)
        self.logger.propagate = False
        for h in list(self.logger.handlers):
            self.logger.removeHandler(h)
        handlers = [logging.StreamHandler(sys.stdout)]
        if to_file:
            path = file_path or "quantization.log"
            os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
            handlers.append(logging.FileHandler(path))
        for h in handlers:
            h.setLevel(lvl)
            h.setFormatter(logging.Formatter(_FORMAT))
            self.logger.addHandler(h)

    @property
    def level(self) -> int:
        return self.logger.level

    def debug(self, msg): self.logger.debug(msg)
    def info(self, msg): self.logger.info(msg)
    def warning(self, msg): self.logger.warning(msg)
    def error(self, msg): self.logger.error(msg)
    def critical(self, msg): self.logger.critical(msg)


def get_logger(name: str = "awq_quantizer", level: str = "INFO", to_file: bool = False,
               file_path: Optional[str] = None) -> Logger:
    return Logger(name=name, level=level, to_file=to_file, file_path=file_path)

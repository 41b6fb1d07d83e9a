"""Quantization package (same re-export as the reference, quantization/__init__.py:5-7)."""

from .awq import AWQQuantizer

__all__ = ["AWQQuantizer"]

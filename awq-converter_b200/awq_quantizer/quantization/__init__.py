"""B200 quantization package.

``AWQQuantizer`` is importable from here exactly as in the reference package
(``from awq_quantizer.quantization import AWQQuantizer``).  Added next to it:

* ``HostArena`` / ``quantize_arena`` -- whole-model flat arena + H2D/K1/D2H pipeline (arena.py)
* ``SearchPipeline`` / ``search_device`` -- activation-aware alpha search on tcgen05 (search.py)
* ``to_autoawq_gemm`` -- AutoAWQ / vLLM checkpoint layout export (export.py)
"""

from .awq import AWQQuantizer
from .arena import HostArena, quantize_arena
from .export import to_autoawq_gemm
from .search import SearchPipeline, search_device

__all__ = ["AWQQuantizer", "HostArena", "quantize_arena", "SearchPipeline", "search_device", "to_autoawq_gemm"]

"""awq_quantizer -- B200-native drop-in for the quantization hot path of shanefitch/AWQ-Converter.

Same import paths as the reference package (``awq_quantizer.quantization.awq.AWQQuantizer``,
``awq_quantizer.utils.tensor_utils.convert_bf16_to_fp16`` ...); the arithmetic runs in the
hand-written sm_100a kernels of ``csrc/`` through the C ABI of ``include/awqk.h``.  There is no CPU
path: without a CUDA device and the built ``libawqk.so`` the compute entry points raise.
"""

__version__ = "0.1.0"

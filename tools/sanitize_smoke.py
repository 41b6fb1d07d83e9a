"""Small invocation of every kernel, for compute-sanitizer (memcheck / racecheck): 
   compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "awq-converter_b200"))
import torch
from awq_quantizer.quantization import AWQQuantizer
from awq_quantizer.utils.tensor_utils import convert_bf16_to_fp16
from tests import datagen
qz = AWQQuantizer(bits=4, group_size=128, symmetric=False, device="cuda:0", logger_level="ERROR", n_grid=3)
w = datagen.weights((72, 1152), "bf16", 1)                    # partial CTA tile for K1 v2 (82944 elements)
r = qz.quantize(w, pack=True)                                 # v1 (unpacked) path
d = qz.dequantize(r)
r2 = qz.quantize_model({"a": w, "b": datagen.weights((40, 384), "bf16", 2), "c": datagen.weights((5, 300), "bf16", 3)}, pack=True,
                       chunk_bytes=1 << 16)                  # v2 via pipeline (flat + row mode) + generic
h = convert_bf16_to_fp16(w)
x = datagen.activations(130, 1152, "bf16", 4)
r3 = qz.quantize(w, activations=x, pack=True)                 # colsum, alpha grid, delta v4, 2-CTA GEMM, K1 col_scale
for g in (32, 64):
    AWQQuantizer(bits=4, group_size=g, symmetric=True, device="cuda:0", logger_level="ERROR").quantize_model({"a": w}, pack=True)
AWQQuantizer(bits=8, group_size=128, symmetric=False, device="cuda:0", logger_level="ERROR").quantize(w, pack=True)
torch.cuda.synchronize()
print("sanitize smoke ok", int(r3["best_idx"]))

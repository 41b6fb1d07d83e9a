#!/usr/bin/env python
"""Full-size check of the activation-aware search (K2) against the oracle's definition: real OPT-350m /
Llama tensor sizes, T = 2048, 20-point grid.  For each tensor the GPU's scale grid is injected into the
oracle ("given equal scales"); reported: max relative error of the 20 error scores, whether the chosen
alpha agrees, and whether the final qweight/qzeros/scales are bit-identical."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "awq-converter_b200"))
import torch
from awq_quantizer.quantization import AWQQuantizer
from awq_quantizer.quantization.search import search_device
from oracle import awq_oracle as O
from tests import datagen

ap = argparse.ArgumentParser()
ap.add_argument("--shapes", default="1024x1024,4096x1024,1024x4096")
ap.add_argument("--tokens", type=int, default=2048)
args = ap.parse_args()
dev = torch.device("cuda:0")
T, n = args.tokens, 20
for spec in args.shapes.split(","):
    C, K = (int(v) for v in spec.split("x"))
    W = datagen.weights((C, K), "bf16", datagen.seed_of("vs", C, K))
    X = datagen.activations(T, K, "bf16", datagen.seed_of("vsx", K))
    qz = AWQQuantizer(bits=4, group_size=128, symmetric=False, device="cuda:0", logger_level="ERROR", n_grid=n)
    t0 = time.time(); got = qz.quantize(W, activations=X, pack=True); t_gpu = time.time() - t0
    r = search_device(W.to(dev), X.to(dev), bits=4, group_size=128, symmetric=False, n_grid=n)
    s_grid = r["s_grid"].cpu()
    t0 = time.time(); want = O.search_scales(W, X, 4, 128, False, n_grid=n, s_grid=s_grid); t_cpu = time.time() - t0
    rel = max(abs(float(got["search_err"][i]) - want["err"][i]) / want["err"][i] for i in range(n))
    own = O.search_scales(W, X, 4, 128, False, n_grid=n)                 # the oracle's own CPU-pow grid
    final = O.pack_result(O.quantize_scaled(W, got["awq_scale"], 4, 128, False))
    exact = all(torch.equal(got[k].view(torch.int16) if got[k].dtype == torch.float16 else got[k],
                            final[k].view(torch.int16) if final[k].dtype == torch.float16 else final[k])
                for k in ("tensor_q", "scales", "zero_points", "qweight", "qzeros"))
    print(json.dumps({"shape": [C, K], "tokens": T, "n_grid": n, "max_rel_err_of_scores": rel,
                      "alpha_gpu": float(got["alpha"]), "alpha_oracle_given_equal_scales": want["alpha"],
                      "alpha_oracle_own_grid": own["alpha"], "final_outputs_bit_exact": exact,
                      "gpu_s_incl_transfers": round(t_gpu, 3), "oracle_s": round(t_cpu, 1)}), flush=True)

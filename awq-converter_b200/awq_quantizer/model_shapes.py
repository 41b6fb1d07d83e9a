"""Tensor inventories of the model shapes BASELINE.json names (synthetic, random-init: there is no
network for checkpoints).  Each entry is (name, shape, calib_key) where calib_key names the
activation tensor an nn.Linear weight would be calibrated with (None = not a linear)."""
from __future__ import annotations

from typing import List, Optional, Tuple

Spec = Tuple[str, tuple, Optional[str]]


def opt_shapes(hidden: int, ffn: int, layers: int, vocab: int = 50272, pos: int = 2050,
               embed_proj: Optional[int] = None, biases: bool = True) -> List[Spec]:
    """facebook/opt-* state-dict layout (decoder only; lm_head is tied to embed_tokens)."""
    e = embed_proj or hidden
    out: List[Spec] = [("model.decoder.embed_tokens.weight", (vocab, e), None),
                       ("model.decoder.embed_positions.weight", (pos, hidden), None)]
    if embed_proj:
        out += [("model.decoder.project_out.weight", (e, hidden), "final"),
                ("model.decoder.project_in.weight", (hidden, e), None)]
    for i in range(layers):
        p = f"model.decoder.layers.{i}."
        for nm in ("k_proj", "v_proj", "q_proj"):
            out.append((p + f"self_attn.{nm}.weight", (hidden, hidden), f"L{i}.attn_in"))
            if biases:
                out.append((p + f"self_attn.{nm}.bias", (hidden,), None))
        out.append((p + "self_attn.out_proj.weight", (hidden, hidden), f"L{i}.attn_out"))
        if biases:
            out.append((p + "self_attn.out_proj.bias", (hidden,), None))
        out += [(p + "self_attn_layer_norm.weight", (hidden,), None), (p + "self_attn_layer_norm.bias", (hidden,), None)]
        out.append((p + "fc1.weight", (ffn, hidden), f"L{i}.mlp_in"))
        out.append((p + "fc2.weight", (hidden, ffn), f"L{i}.mlp_mid"))
        if biases:
            out += [(p + "fc1.bias", (ffn,), None), (p + "fc2.bias", (hidden,), None)]
        out += [(p + "final_layer_norm.weight", (hidden,), None), (p + "final_layer_norm.bias", (hidden,), None)]
    return out


def llama_shapes(hidden: int, ffn: int, layers: int, heads: int, kv_heads: int, vocab: int = 128256) -> List[Spec]:
    hd = hidden // heads
    kv = kv_heads * hd
    out: List[Spec] = [("model.embed_tokens.weight", (vocab, hidden), None)]
    for i in range(layers):
        p = f"model.layers.{i}."
        out += [(p + "self_attn.q_proj.weight", (hidden, hidden), f"L{i}.attn_in"),
                (p + "self_attn.k_proj.weight", (kv, hidden), f"L{i}.attn_in"),
                (p + "self_attn.v_proj.weight", (kv, hidden), f"L{i}.attn_in"),
                (p + "self_attn.o_proj.weight", (hidden, hidden), f"L{i}.attn_out"),
                (p + "mlp.gate_proj.weight", (ffn, hidden), f"L{i}.mlp_in"),
                (p + "mlp.up_proj.weight", (ffn, hidden), f"L{i}.mlp_in"),
                (p + "mlp.down_proj.weight", (hidden, ffn), f"L{i}.mlp_mid"),
                (p + "input_layernorm.weight", (hidden,), None),
                (p + "post_attention_layernorm.weight", (hidden,), None)]
    out += [("model.norm.weight", (hidden,), None), ("lm_head.weight", (vocab, hidden), None)]
    return out


WORKLOADS = {
    # BASELINE.json configs[0]: exactly the tensors of the reference's test_quantization.py:54-63
    "test_quantization": lambda: [("layer1.weight", (768, 3072), None), ("layer2.weight", (768, 3, 768), None),
                                  ("small.weight", (10, 10), None)],
    "opt-125m": lambda: opt_shapes(768, 3072, 12),
    "opt-350m": lambda: opt_shapes(1024, 4096, 24, embed_proj=512),          # configs[1]
    "llama3-8b": lambda: llama_shapes(4096, 14336, 32, 32, 8),               # configs[2]
    "llama3-70b": lambda: llama_shapes(8192, 28672, 80, 64, 8),              # configs[3]
    "micro-8192x28672": lambda: [("w", (8192, 28672), None)],                # configs[4]
}


def workload(name: str) -> List[Spec]:
    return WORKLOADS[name]()


def numel(shape) -> int:
    n = 1
    for s in shape:
        n *= s
    return n


def total_params(specs: List[Spec]) -> int:
    return sum(numel(s) for _, s, _ in specs)


def partition_lpt(items: List[Tuple[str, int]], parts: int) -> List[List[str]]:
    """Greedy largest-first onto the least-loaded bin -- the algorithm of the reference's (never
    called) partition_tensors, main.py:395-427.  Deterministic: ties go to the lowest bin index and
    equal sizes keep their input order."""
    if parts <= 1:
        return [[n for n, _ in items]]
    order = sorted(range(len(items)), key=lambda i: (-items[i][1], i))
    bins: List[List[str]] = [[] for _ in range(parts)]
    load = [0] * parts
    for i in order:
        j = load.index(min(load))
        bins[j].append(items[i][0])
        load[j] += items[i][1]
    return bins

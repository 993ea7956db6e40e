"""Deterministic parameter / input generators shared by make_golden.py and the tests.

Fixtures under tests/golden/*.npz hold only reference OUTPUTS (plus inputs that come
from the reference sampler).  Parameters are regenerated from a seed by ``make_params``
so the fixtures stay small; make_golden.py loads exactly these tensors into the
reference model with ``load_state_dict`` before running it.
"""
from __future__ import annotations

import os

import numpy as np
import torch

GOLDEN_DIR = os.path.dirname(os.path.abspath(__file__))


def make_params(seed: int, V: int, d: int, L: int, hid: int, bs: int, *, isInC=False, isItC=True,
                isDR=False, zero_rows=(), zero_pos=()) -> dict:
    """Reference-named state dict with non-trivial values (LN gains != 1, biases != 0).

    L is the *encoder* length (already doubled by the caller when isInC, model_seq.py:399-400).
    """
    rng = np.random.default_rng(seed)

    def U(shape, fan_in):
        b = 1.0 / np.sqrt(fan_in)
        return torch.from_numpy(rng.uniform(-b, b, size=shape).astype(np.float32))

    def N(shape, s=1.0, mu=0.0):
        return torch.from_numpy((mu + s * rng.standard_normal(size=shape)).astype(np.float32))

    P = {"item_emb_layer.emb_item.weight": N((V, d))}
    for r in zero_rows:
        P["item_emb_layer.emb_item.weight"][r] = 0.0
    mims = (["inc_d1", "inc_d2"] if isInC else []) + (["itc_d1", "itc_d2"] if isItC else [])
    for m in mims:
        P[f"{m}.trans_nn.weight"] = U((d, d), d)
        P[f"{m}.trans_nn.bias"] = U((d,), d)
        P[f"{m}.trans_bs.weight"] = U((1, bs), bs)
        P[f"{m}.trans_bs.bias"] = U((1,), bs)
    for s in ("sac1", "sac2"):
        P[f"{s}.pos_emb.weight"] = N((L, d))
        for r in zero_pos:
            P[f"{s}.pos_emb.weight"][r] = 0.0
        P[f"{s}.last_layernorm.weight"] = N((d,), 0.1, 1.0)
        P[f"{s}.last_layernorm.bias"] = N((d,), 0.1)
        for i in range(2):
            P[f"{s}.attention_layernorms.{i}.weight"] = N((d,), 0.1, 1.0)
            P[f"{s}.attention_layernorms.{i}.bias"] = N((d,), 0.1)
            P[f"{s}.attention_layers.{i}.in_proj_weight"] = U((3 * d, d), d / 3.0)
            P[f"{s}.attention_layers.{i}.in_proj_bias"] = N((3 * d,), 0.05)
            P[f"{s}.attention_layers.{i}.out_proj.weight"] = U((d, d), d)
            P[f"{s}.attention_layers.{i}.out_proj.bias"] = N((d,), 0.05)
            P[f"{s}.forward_layernorms.{i}.weight"] = N((d,), 0.1, 1.0)
            P[f"{s}.forward_layernorms.{i}.bias"] = N((d,), 0.1)
            for c in ("conv1", "conv2"):
                P[f"{s}.forward_layers.{i}.{c}.weight"] = U((d, d, 1), d)
                P[f"{s}.forward_layers.{i}.{c}.bias"] = U((d,), d)
    heads = ["predictModule"] + (["predict_ips", "predict_gfunc"] if isDR else [])
    for h in heads:
        P[f"{h}.fc.0.weight"] = U((hid, 2 * d), 2 * d)
        P[f"{h}.fc.0.bias"] = U((hid,), 2 * d)
        P[f"{h}.fc.2.weight"] = U((1, hid), hid)
        P[f"{h}.fc.2.bias"] = U((1,), hid)
    return P


def make_keep_masks(seed: int, B: int, L: int, d: int, heads: int = 8, p: float = 0.5) -> dict:
    """Boolean dropout keep-masks for both encoders, in the oracle's layout."""
    g = torch.Generator().manual_seed(seed)

    def bern(shape):
        return torch.rand(shape, generator=g) >= p

    out = {}
    for s in ("sac1", "sac2"):
        m = {"emb": bern((B, L, d))}
        for i in range(2):
            m[f"attn{i}"] = bern((B, heads, L, L))
            m[f"ffn1_{i}"] = bern((B, L, d))
            m[f"ffn2_{i}"] = bern((B, L, d))
        out[s] = m
    return out


def load(name: str) -> dict:
    with np.load(os.path.join(GOLDEN_DIR, name), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}

"""Same-box GPU-eager incumbent (last row): the oracle -- a plain-torch restatement of the
reference path -- executed with stock ATen kernels on cuda:0 at the bench workload, with the closed-form ItC
(the literal one needs 275 GB at C3) and torch.optim.Adam over all parameters including the dense table.

This is a measurement script, not a test and not part of the product: it lives under tests/ because only
tests/ may execute oracle/.  Usage:  python tests/eager_gpu_baseline.py [--batch 1024 --seq-len 200 --steps 10]
Prints one JSON line per matmul mode (fp32 and allow_tf32)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from common import make_params  # noqa: E402
from oracle import amid_oracle as O  # noqa: E402

V, D, HID = 894_820, 128, 32


def masks_on_device(B, L, dev, p=0.5):
    def bern(shape):
        return torch.rand(shape, device=dev) >= p
    out = {}
    for s in ("sac1", "sac2"):
        m = {"emb": bern((B, L, D))}
        for i in range(2):
            m[f"attn{i}"] = bern((B, 8, L, L))
            m[f"ffn1_{i}"] = bern((B, L, D))
            m[f"ffn2_{i}"] = bern((B, L, D))
        out[s] = m
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--seq-len", type=int, default=200)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    B, L, C = a.batch, a.seq_len, 2
    rng = np.random.default_rng(0)
    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        P = {k: v.to(dev).requires_grad_(True) for k, v in make_params(7, V, D, L, HID, B, isDR=False).items()}
        opt = torch.optim.Adam(list(P.values()), lr=5e-4)
        batches = []
        for _ in range(4):
            batches.append({
                "i_node": torch.from_numpy(rng.integers(0, V, B)).to(dev),
                "neg": torch.from_numpy(rng.integers(0, V, (B, C - 1))).to(dev),
                "s1": torch.from_numpy(rng.integers(0, V, (B, L))).to(dev),
                "s2": torch.from_numpy(rng.integers(0, V, (B, L))).to(dev),
                "label": torch.tensor([[1.0, 0.0]], device=dev).repeat(B, 1),
                "dom": torch.from_numpy(rng.integers(0, 2, B)).to(dev),
            })

        def step(i):
            b = batches[i % 4]
            masks = masks_on_device(B, L, dev)
            outs = O.sasrec_forward(P, b["i_node"], b["neg"], b["s1"], b["s2"], isInC=False, isItC=True, ts1=0.5,
                                    ts2=0.4, isDR=False, masks=masks, closed_form=True)
            loss = O.loss_cls(outs[0], outs[1], b["label"], b["dom"])
            opt.zero_grad()
            loss.backward()
            opt.step()
            return loss

        for i in range(a.warmup):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.steps):
            loss = step(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        print(json.dumps({"impl": "gpu-eager oracle (stock ATen kernels)", "matmul": "tf32" if tf32 else "fp32",
                          "metric": "train_seqs_per_sec", "value": B / (ms / 1e3), "ms_per_step": ms, "batch": B,
                          "seq_len": L, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30,
                          "loss": float(loss)}), flush=True)
        del P, opt
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

"""Import the UNMODIFIED reference modules from oracle/_ref (see oracle/build_ref.py) -- test / benchmark
infrastructure only; nothing under amid_b200/ may import this.

Shims, applied by monkey-patching at import time, never by editing the files:
  1. `random.sample` on a set raises TypeError on Python >= 3.11 (dataset_seq.py:198,201,216,219,512,...):
     sets are converted to tuples first, which is what Python <= 3.10 did implicitly.
  2. device="cpu" only, and only inside ``with ref.on_cpu():`` -- the hard-coded `.cuda()` / `device="cuda"`
     (model_seq.py:362,365,369; train_sr.py:191-207) become no-ops so that the path runs on the host cores
     (BASELINE config 1: "device forced to CPU"); the patch is undone when the block exits.

`closed_form_itc(ms)` additionally swaps InterComp.forward (model_seq.py:483-497) for its exact closed form
(max deviation 6e-8 from the literal code, tests/golden/mim_peaked.npz): the literal code
materialises [bs, bs, n, n] tensors -- 275 GB at the C3 shape -- so it cannot execute there.  Every record produced
with the patch says so.
"""
from __future__ import annotations

import importlib
import os
import random
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
_orig_sample = random.sample


def _sample(pop, k, **kw):
    if isinstance(pop, (set, frozenset)):
        pop = tuple(pop)
    return _orig_sample(pop, k, **kw)


def available() -> bool:
    return os.path.exists(os.path.join(REF, "model_seq.py"))


class _CpuShims:
    """Shim 2 as a context manager: while active, `.cuda()` / `device="cuda"` are no-ops so that the reference's
    forward (model_seq.py:362,365,369) runs on the host cores; everything is restored on exit, so the GPU arm of
    the same process is never affected."""

    def __enter__(self):
        import torch
        import torch.nn as nn
        self._saved = (torch.Tensor.cuda, nn.Module.cuda, torch.ones)
        _ones = torch.ones

        def ones(*a, **k):
            k.pop("device", None)
            return _ones(*a, **k)

        torch.Tensor.cuda = lambda self, *a, **k: self
        nn.Module.cuda = lambda self, *a, **k: self
        torch.ones = ones
        return self

    def __exit__(self, *exc):
        import torch
        import torch.nn as nn
        torch.Tensor.cuda, nn.Module.cuda, torch.ones = self._saved
        return False


def load(device: str = "cpu") -> types.SimpleNamespace:
    """Returns a namespace with the reference's model_seq, dataset_seq and utils modules.  With device="cpu" every
    call into the reference must run inside ``with ref.on_cpu():`` (shim 2); shim 1 is installed permanently."""
    if not available():
        raise FileNotFoundError(f"{REF} is missing: run `python oracle/build_ref.py` where /root/reference exists")
    random.sample = _sample
    mods = {}
    saved = {n: sys.modules.pop(n) for n in ("model_seq", "dataset_seq", "utils") if n in sys.modules}
    sys.path.insert(0, REF)
    try:
        for n in ("utils", "model_seq", "dataset_seq"):
            mods[n] = importlib.import_module(n)
    finally:
        sys.path.remove(REF)
        for n in ("model_seq", "dataset_seq", "utils"):
            sys.modules.pop(n, None)
        sys.modules.update(saved)
    mods["dataset_seq"].random.sample = _sample
    import contextlib
    return types.SimpleNamespace(model_seq=mods["model_seq"], dataset_seq=mods["dataset_seq"], utils=mods["utils"],
                                 data_root=os.path.join(REF, "amazon_dataset"),
                                 on_cpu=_CpuShims if device == "cpu" else contextlib.nullcontext)


def closed_form_itc(ms) -> None:
    """InterComp.forward -> closed form (same parameters, same output, O(bs n d) memory)."""
    import torch

    def forward(self, seq_d1, seq_d2):
        # m_j = max_{s,t} <seq_d1[j,s], seq_d2[j,t]>; p = softmax_j(m); g_j = [p_j > ts]            (:487-492)
        m = torch.einsum("jsd,jtd->jst", seq_d1, seq_d2).flatten(1).max(1)[0]
        p = torch.softmax(m, dim=0)
        g = (p > self.threshold).to(seq_d2.dtype).detach()
        # Y_j = (g_j seq_d2[j]) W_nn^T + b_nn ; E = sum_j w_bs[j] Y_j + b_bs                          (:493-495)
        y = self.trans_nn(seq_d2 * g.view(-1, 1, 1))
        e = torch.einsum("j,jnd->nd", self.trans_bs.weight.view(-1), y) + self.trans_bs.bias
        return torch.cat((seq_d1, e.unsqueeze(0).expand(seq_d1.shape[0], -1, -1)), dim=1)           # (:496-497)

    ms.InterComp.forward = forward

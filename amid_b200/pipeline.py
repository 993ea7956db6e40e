"""Device-side input pipeline replacing DualDomainSeqDataset.__getitem__ + collate_fn_enhance
(dataset_seq.py:137-274) on the step path.

The reference re-parses two JSON histories, rebuilds a Python set and calls random.sample for every row of every
epoch in DataLoader workers, then ships ten float32 tensors per batch to the GPU field by field
(train_sr.py:191-200).  Here the histories are tokenised ONCE on the host (``prepare_rows``) into CSR arrays that
live in HBM, and a whole batch -- padded histories, targets, flags and (optionally) freshly sampled negatives --
is produced by one kernel launch (csrc/pipeline.cu) directly as int64 device tensors.

Two modes:
  * ``batch(rows, negatives=...)``: "replay" -- negatives drawn elsewhere (e.g. by the reference sampler) are
    used as given; every field is then bit-identical to the reference's collated batch (tests/test_gpu_pipeline.py).
    Evaluation PARITY with the reference protocol requires this mode.
  * ``batch(rows, k=..., seed=...)``: negatives are drawn on the device: k distinct items of the target domain's
    pool outside the user's full own-domain sequence, reproducible in (seed, row).  The draw is the first k admissible items of a
    keyed pseudo-random PERMUTATION of the pool (4-round Feistel network with cycle walking, keyed by (seed, row)): a
    duplicate-free pseudo-random k-subset, the role random.sample plays in the reference sampler.  The stream differs from
    Python's, so bit-for-bit evaluation parity with the reference protocol still requires replay mode.
"""
from __future__ import annotations

import ctypes as C
import json
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch

from . import _abi
from ._abi import BatchOut, BatchSource, call


def _csr(lists: Sequence[Sequence[int]]):
    offs = np.zeros(len(lists) + 1, dtype=np.int64)
    np.cumsum([len(x) for x in lists], out=offs[1:])
    vals = np.fromiter((v for x in lists for v in x), dtype=np.int64, count=int(offs[-1]))
    return vals, offs


def prepare_rows(user_ids: Sequence[int], seq_d1: Sequence[Sequence[int]], seq_d2: Sequence[Sequence[int]],
                 domain_id: Sequence[int], ob_label: Optional[Sequence[int]] = None) -> Dict[str, np.ndarray]:
    """One pass over the table (host, once per dataset): the deterministic part of __getitem__.
    For the row's own domain the target is the last item and it is removed from the history together with all its
    earlier occurrences (dataset_seq.py:189-195 / 207-213); the other domain's history is used whole; negatives must
    avoid the FULL own-domain sequence (:188 / :206); the item pools are the sets of all items per domain (:151-158)."""
    n = len(user_ids)
    if not (len(seq_d1) == len(seq_d2) == len(domain_id) == n):
        raise ValueError("prepare_rows: column lengths differ")
    h1, h2, ex = [], [], []
    target = np.zeros(n, dtype=np.int64)
    overlap = np.zeros(n, dtype=np.int32)
    domain = np.zeros(n, dtype=np.int32)
    pool1, pool2 = set(), set()
    for i in range(n):
        s1, s2 = list(seq_d1[i]), list(seq_d2[i])
        pool1.update(s1)
        pool2.update(s2)
        overlap[i] = 1 if (s1 and s2) else 0
        dom = 0 if int(domain_id[i]) == 0 else 1
        domain[i] = dom
        own = s1 if dom == 0 else s2
        if not own:
            raise ValueError(f"row {i}: empty own-domain sequence (the reference would raise IndexError here)")
        item = own[-1]
        target[i] = item
        hist = [x for x in own[:-1] if x != item]
        ex.append(sorted(set(own)))
        h1.append(hist if dom == 0 else s1)
        h2.append(hist if dom == 1 else s2)
    out = {"user": np.asarray(user_ids, dtype=np.int64), "target": target, "domain": domain, "overlap": overlap,
           "pool_d1": np.asarray(sorted(pool1), dtype=np.int64), "pool_d2": np.asarray(sorted(pool2), dtype=np.int64)}
    for name, lists in (("hist_d1", h1), ("hist_d2", h2), ("excl", ex)):
        out[name + "_vals"], out[name + "_offs"] = _csr(lists)
    if ob_label is not None:                       # DualDomainSeqDatasetDR (dataset_seq.py:453, 498, 551): copied per row
        if len(ob_label) != n:
            raise ValueError("prepare_rows: ob_label length differs")
        out["ob_label"] = np.asarray(ob_label, dtype=np.int64)
    return out


def prepare_csv(csv_path: str) -> Dict[str, np.ndarray]:
    """The reference's CSV layout (user_id, seq_d1, seq_d2 as JSON lists, domain_id; dataset_seq.py:140-146)."""
    import pandas as pd
    df = pd.read_csv(csv_path)
    return prepare_rows(df["user_id"].tolist(), [json.loads(s) for s in df["seq_d1"]], [json.loads(s) for s in df["seq_d2"]],
                        df["domain_id"].tolist(), df["ob_label"].tolist() if "ob_label" in df.columns else None)


class DeviceDataset:
    """The prepared table resident in HBM; ``batch`` builds one training / eval batch on the device."""

    def __init__(self, prep: Dict[str, np.ndarray], seq_len: int, long_length: int, pad_id: int, device="cuda"):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _abi.AmidError("DeviceDataset needs a CUDA device (no CPU fallback)")
        self.seq_len, self.long_length, self.pad_id = int(seq_len), int(long_length), int(pad_id)
        self.t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in prep.items()}
        self.n_rows = int(prep["target"].shape[0])
        self.pool_d1, self.pool_d2 = self.t["pool_d1"], self.t["pool_d2"]
        t = self.t
        # zero-length CSR value arrays still need a valid pointer
        for k in ("hist_d1_vals", "hist_d2_vals", "excl_vals"):
            if t[k].numel() == 0:
                t[k] = torch.zeros(1, device=dev, dtype=torch.int64)
        self.src = BatchSource(t["hist_d1_vals"].data_ptr(), t["hist_d1_offs"].data_ptr(), t["hist_d2_vals"].data_ptr(),
                               t["hist_d2_offs"].data_ptr(), t["excl_vals"].data_ptr(), t["excl_offs"].data_ptr(),
                               t["target"].data_ptr(), t["user"].data_ptr(), t["domain"].data_ptr(), t["overlap"].data_ptr(),
                               t["pool_d1"].data_ptr(), t["pool_d2"].data_ptr(), t["pool_d1"].numel(), t["pool_d2"].numel(),
                               self.n_rows)
        self.dev = dev

    def __len__(self) -> int:
        return self.n_rows

    def batch(self, rows: torch.Tensor, negatives: Optional[torch.Tensor] = None, k: int = 1, seed: int = 0,
              check: bool = False) -> Dict[str, torch.Tensor]:
        """rows: int64 [B] dataset row per batch position (any order, repeats allowed).  Returns the dict the drivers
        build from the reference's collate (train_sr.py:191-200), ids as int64 device tensors, label as fp32."""
        from .hotpath import _ptr, _stream
        rows = rows.to(self.dev, torch.int64).contiguous()
        B, L = rows.numel(), self.seq_len
        i64 = lambda *s: torch.empty(*s, device=self.dev, dtype=torch.int64)
        out = {"seq_d1": i64(B, L), "seq_d2": i64(B, L), "i_node": i64(B), "user_node": i64(B), "domain_id": i64(B),
               "overlap_label": i64(B), "long_tail_mask_d1": i64(B), "long_tail_mask_d2": i64(B)}
        if negatives is not None:
            neg = negatives.to(self.dev, torch.int64).reshape(B, -1).contiguous()
            K, sample = neg.shape[1], 0
        else:
            K, sample = int(k), int(k)
            neg = torch.zeros(B, K, device=self.dev, dtype=torch.int64)   # a row whose pool is exhausted keeps id 0 + error flag
        o = BatchOut(*[out[n].data_ptr() for n in ("seq_d1", "seq_d2", "i_node", "user_node", "domain_id", "overlap_label",
                                                   "long_tail_mask_d1", "long_tail_mask_d2")], neg.data_ptr() if sample else None)
        call("amid_batch_build", C.byref(self.src), _ptr(rows), B, L, sample, self.long_length, self.pad_id,
             int(seed) & (2**64 - 1), C.byref(o), _stream())
        if check:
            code = _abi.lib().amid_gather_error_host_sync()
            if code:
                raise IndexError("batch: row index out of range or item pool exhausted by the exclusion list")
        out["neg_samples"] = neg
        if "ob_label" in self.t:                   # the DR drivers' extra field (train_sr_dr.py:201, 374)
            out["ob_label"] = self.t["ob_label"][rows]
        out["label"] = torch.cat((torch.ones(B, 1, device=self.dev), torch.zeros(B, K, device=self.dev)), 1)
        return out

    def epoch(self, batch_size: int, k: int = 1, seed: int = 0, shuffle: bool = True, drop_last: bool = True,
              rank: int = 0, world: int = 1) -> Iterable[Dict]:
        """Batches of one epoch (DataLoader(shuffle=True, drop_last=True) of train_sr.py:452-455), built on the device.
        ``batch_size`` is the GLOBAL batch; under data parallelism rank r gets rows [r*B/world, (r+1)*B/world) of every
        global batch (the slice DistCtx assigns it).  The sampler seed of a batch is a 64-bit mix of (seed, batch index),
        so epochs never share a negative stream; the first batch of an epoch is range-checked on the host."""
        if batch_size % world:
            raise ValueError(f"global batch {batch_size} is not divisible by the world size {world}")
        g = torch.Generator(device="cpu").manual_seed(int(seed))
        order = torch.randperm(self.n_rows, generator=g) if shuffle else torch.arange(self.n_rows)
        order = order.to(self.dev)
        stop = self.n_rows - (self.n_rows % batch_size if drop_last else 0)
        bl = batch_size // world
        for bi, i in enumerate(range(0, stop, batch_size)):
            rows = order[i:i + batch_size]
            if world > 1:
                rows = rows[rank * bl:(rank + 1) * bl]
            yield self.batch(rows, k=k, seed=_mix64(int(seed), bi), check=(bi == 0))


def _mix64(seed: int, index: int) -> int:
    """splitmix64 of the pair: distinct (seed, batch index) pairs give unrelated sampler seeds."""
    z = ((seed & (2**64 - 1)) * 0x9E3779B97F4A7C15 + (index + 1) * 0xD1B54A32D192ED03) & (2**64 - 1)
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2**64 - 1)
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2**64 - 1)
    return z ^ (z >> 31)

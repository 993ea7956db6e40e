"""Evaluation ranking on the device (reference: test() in train_sr.py:31-128 and
utils.py:21-68, 296-313).

The scores stay in HBM; one kernel counts, for every user row, how many candidates beat
(and how many tie with) the positive in column 0.  rank = n_greater when there is no tie;
rows with ties are resolved on the host with the *same numpy expression the reference
uses* (utils.py:297), because numpy's argsort order among equal keys is implementation
defined and the requirement is bit-exact rankings.  Metrics are accumulated in float64 in
row order, exactly like the reference's Python loop (utils.py:303-313).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from ._abi import call
from .hotpath import _ptr, _stream

FIX_VALUE = 1e-7   # train_sr.py:42


def rank_of_positive(scores: torch.Tensor, fix: float = 0.0) -> np.ndarray:
    """rank = argsort(argsort(-scores))[:, 0] with scores[:,0] -= fix applied first (fp32)."""
    if scores.dim() != 2 or scores.dtype != torch.float32 or not scores.is_cuda:
        raise ValueError("scores must be a CUDA float32 [N, C] tensor")
    scores = scores.contiguous()
    N, C = scores.shape
    if N == 0:
        return np.zeros(0, dtype=np.int64)
    ng = torch.empty(N, device=scores.device, dtype=torch.int32)
    ne = torch.empty(N, device=scores.device, dtype=torch.int32)
    call("amid_rank_counts", _ptr(scores), N, C, float(np.float32(fix)), _ptr(ng), _ptr(ne), _stream())
    both = torch.stack((ng, ne)).cpu().numpy()
    ranks = both[0].astype(np.int64)
    tied = np.nonzero(both[1])[0]
    if len(tied):
        rows = scores[torch.from_numpy(tied).to(scores.device)].cpu().numpy()
        rows[:, 0] = rows[:, 0] - fix                     # same float32 arithmetic as train_sr.py:114
        ranks[tied] = (-rows).argsort().argsort()[:, 0]   # utils.py:297
    return ranks


def metrics_from_ranks(ranks: np.ndarray):
    """(HIT@1, NDCG@1, HIT@5, NDCG@5, HIT@10, NDCG@10, MRR): utils.py:296-313, float64, row order."""
    n = len(ranks)
    r = ranks.astype(np.float64)
    out = []
    for k in (1, 5, 10):
        hit = float(np.count_nonzero(ranks < k))
        terms = np.where(ranks < k, 1.0 / np.log2(r + 2.0), 0.0)
        ndcg = float(np.cumsum(terms)[-1]) if n else 0.0   # cumsum = the reference's sequential +=
        out += [hit / n, ndcg / n]
    mrr = float(np.cumsum(1.0 / (r + 1.0))[-1]) if n else 0.0
    out.append(mrr / n)
    return tuple(out)


def evaluate_lists(pred_d1: torch.Tensor, pred_d2: torch.Tensor, domain_id: torch.Tensor,
                   overlap_label: Optional[torch.Tensor] = None) -> Dict[str, tuple]:
    """The list bookkeeping of test(): rows with domain_id == 0 are scored with pred_d1, the
    others with pred_d2 (utils.py:21-32); aggregate lists get the 1e-7 fix on the positive
    (train_sr.py:114-115 / 124-125), the overlap / non-overlap lists do not (:120-123)."""
    res = {}
    is1 = domain_id == 0
    for name, pred, sel in (("d1", pred_d1, is1), ("d2", pred_d2, ~is1)):
        if overlap_label is not None:
            for tag, osel in (("ov", overlap_label != 0), ("no", overlap_label == 0)):
                rows = pred[sel & osel]
                if rows.shape[0]:
                    res[f"{name}_{tag}"] = metrics_from_ranks(rank_of_positive(rows, 0.0))
        rows = pred[sel]
        if rows.shape[0]:
            res[name] = metrics_from_ranks(rank_of_positive(rows, FIX_VALUE))
    return res

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


# --------------------------------------------------------------------------------------------------
# Full-catalogue evaluation (BASELINE config 5): every user against the whole target-domain pool.
# The reference only ever scores 1 + neg_nums sampled candidates (dataset_seq.py:201); here the candidate
# list of a user is [positive, every other pool item in pool order], scored by the same predictModule and
# ranked by the same rule.  The U x I score matrix is never materialised: csrc/catalogue.cu counts, per
# user, the pool items that beat / tie the positive.
# --------------------------------------------------------------------------------------------------
class Catalogue:
    """Item halves ``Bc = W0[:, 128:] item + b0`` of predictModule for the item pools of both domains.
    They do not depend on the users, so they are built once per evaluation (call ``refresh`` after the
    weights change).  ``pool_d1`` / ``pool_d2`` are int64 item ids (dataset_seq.py:141-142, 151-158)."""

    def __init__(self, P: Dict[str, torch.Tensor], cfg, pool_d1: torch.Tensor, pool_d2: torch.Tensor):
        dev = P["item_emb_layer.emb_item.weight"].device
        if cfg.hid_dim != 32:
            raise ValueError("the full-catalogue path is written for hid_dim == 32")
        self.P, self.cfg = P, cfg
        self.pools = [pool_d1.to(dev, torch.int64).contiguous(), pool_d2.to(dev, torch.int64).contiguous()]
        n1, n2 = self.pools[0].numel(), self.pools[1].numel()
        if n1 == 0 or n2 == 0:
            raise ValueError("empty item pool")
        self.ranges = [(0, n1), (n1, n1 + n2)]
        V = P["item_emb_layer.emb_item.weight"].shape[0]
        self.index_of = []                         # per domain: item id -> row of Bc (-1 = not in the pool)
        for d, (lo, hi) in enumerate(self.ranges):
            m = torch.full((V,), -1, device=dev, dtype=torch.int32)
            m[self.pools[d]] = torch.arange(lo, hi, device=dev, dtype=torch.int32)
            self.index_of.append(m)
        self.Bc = torch.empty(n1 + n2, 32, device=dev, dtype=torch.float32)
        self.refresh()

    def refresh(self):
        P = self.P
        table = P["item_emb_layer.emb_item.weight"]
        ids = torch.cat(self.pools)
        call("amid_catalogue_item_proj", _ptr(table), table.shape[0], _ptr(ids), ids.numel(),
             _ptr(P["predictModule.fc.0.weight"]), _ptr(P["predictModule.fc.0.bias"]), 32, _ptr(self.Bc), _stream())
        if _abi_gather_error():
            raise IndexError("catalogue: item id out of range")


def _abi_gather_error() -> bool:
    from ._abi import lib
    return bool(lib().amid_gather_error_host_sync())


@torch.no_grad()
def full_catalogue_ranks(P: Dict[str, torch.Tensor], cfg, cat: Catalogue, batch: Dict[str, torch.Tensor], dist=None):
    """Ranks of the positives of one eval batch against the whole pool of their target domain.
    Returns {dom: (user_rows, ranks_fix, ranks_nofix)} with numpy int64 arrays; ranks_fix applies the 1e-7 fix
    (aggregate lists, train_sr.py:114-115), ranks_nofix does not (overlap / non-overlap lists, :120-123)."""
    from . import hotpath
    i_node = batch["i_node"]
    dev = i_node.device
    B = i_node.shape[0]
    neg = batch["neg_samples"][:, :1].contiguous() if "neg_samples" in batch else i_node.view(B, 1).clone()
    _, ctx = hotpath.forward(P, cfg, i_node, neg, batch["seq_d1"], batch["seq_d2"], train=False, seed=0, dist=dist,
                             need_ctx=True)
    A = torch.empty(B, 2, 32, device=dev, dtype=torch.float32)
    s = _stream()
    call("amid_catalogue_user_proj", _ptr(ctx.us[0]), _ptr(ctx.us[1]), B, _ptr(P["predictModule.fc.0.weight"]), 32, _ptr(A), s)
    w2, b2 = P["predictModule.fc.2.weight"], P["predictModule.fc.2.bias"]
    out = {}
    fix32 = float(np.float32(FIX_VALUE))
    for dom in (0, 1):
        rows = torch.nonzero(batch["domain_id"] == dom).flatten().to(torch.int32)
        n = rows.numel()
        if n == 0:
            continue
        pos_idx = cat.index_of[dom][i_node.reshape(-1)].contiguous()
        if bool((pos_idx[rows.long()] < 0).any()):
            raise IndexError(f"full_catalogue_ranks: a positive item of domain {dom + 1} is not in its pool")
        lo, hi = cat.ranges[dom]
        counts = torch.empty(n, 4, device=dev, dtype=torch.int32)
        s_pos = torch.empty(n, device=dev, dtype=torch.float32)
        call("amid_catalogue_rank", _ptr(A), _ptr(rows), n, dom, _ptr(cat.Bc), lo, hi, _ptr(pos_idx), _ptr(w2), _ptr(b2),
             fix32, _ptr(counts), _ptr(s_pos), s)
        c = counts.cpu().numpy().astype(np.int64)
        ranks_nofix, ranks_fix = c[:, 0].copy(), c[:, 2].copy()
        tied = np.nonzero((c[:, 1] > 0) | (c[:, 3] > 0))[0]
        if len(tied):                              # resolve with the reference's own numpy expression (utils.py:297)
            sel = rows[torch.from_numpy(tied).to(dev)].contiguous()
            sc = torch.empty(len(tied), hi - lo, device=dev, dtype=torch.float32)
            sp = torch.empty(len(tied), device=dev, dtype=torch.float32)
            call("amid_catalogue_scores", _ptr(A), _ptr(sel), len(tied), dom, _ptr(cat.Bc), lo, hi, _ptr(pos_idx), _ptr(w2),
                 _ptr(b2), _ptr(sp), _ptr(sc), s)
            sc, sp = sc.cpu().numpy(), sp.cpu().numpy()
            pcol = (pos_idx[sel.long()] - lo).cpu().numpy()
            for k, t in enumerate(tied):
                others = np.delete(sc[k], pcol[k]) if 0 <= pcol[k] < hi - lo else sc[k]
                for fixv, dst in ((0.0, ranks_nofix), (FIX_VALUE, ranks_fix)):
                    row = np.concatenate((np.array([sp[k]], dtype=np.float32), others))
                    row[0] = row[0] - fixv
                    dst[t] = (-row).argsort().argsort()[0]
        out[dom] = (rows.cpu().numpy().astype(np.int64), ranks_fix, ranks_nofix)
    return out


def evaluate_full_catalogue(P, cfg, cat: Catalogue, batches, dist=None) -> Dict[str, tuple]:
    """test() of train_sr.py:31-128 with the whole pool as the candidate list: the same six lists
    (d1, d2, and their overlap / non-overlap splits) and the same seven metrics per list.  With ``dist`` the
    caller shards whole batches across ranks; rank lists are gathered in rank order before the metrics."""
    acc = {k: [] for k in ("d1", "d2", "d1_ov", "d1_no", "d2_ov", "d2_no")}
    for b in batches:
        r = full_catalogue_ranks(P, cfg, cat, b, dist=None)
        ov = b["overlap_label"].cpu().numpy() if "overlap_label" in b else None
        for dom, name in ((0, "d1"), (1, "d2")):
            if dom not in r:
                continue
            rows, rf, rn = r[dom]
            acc[name].append(rf)
            if ov is not None:
                o = ov[rows] != 0
                acc[name + "_ov"].append(rn[o])
                acc[name + "_no"].append(rn[~o])
    res = {}
    for k, parts in acc.items():
        ranks = np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64)
        if dist is not None and dist.world > 1:
            ranks = _gather_ranks(ranks, dist, batches[0]["i_node"].device if len(batches) else torch.device("cuda"))
        if len(ranks):
            res[k] = metrics_from_ranks(ranks)
    return res


# --------------------------------------------------------------------------------------------------
# Launch-bound eval batches: CUDA-graph replay of the eval-mode forward (fixed shapes, static input buffers)
# --------------------------------------------------------------------------------------------------
class GraphedForward:
    """The eval-mode forward of one batch shape captured ONCE in a CUDA graph (C1-sized batches are ~100 launches of a
    few microseconds each: launch-bound).  ``run(batch)`` copies the ids into the static input buffers and replays the
    graph; the returned tensors are the graph's static outputs (valid until the next ``run``).  Dropout is off in
    eval mode, so nothing step-dependent is baked into the graph."""

    def __init__(self, P, cfg, B: int, L: int, C: int, with_user_proj: bool = False):
        from . import hotpath
        dev = P["item_emb_layer.emb_item.weight"].device
        i64 = lambda *s_: torch.zeros(*s_, device=dev, dtype=torch.int64)
        self.inp = {"i_node": i64(B), "neg_samples": i64(B, C - 1), "seq_d1": i64(B, L), "seq_d2": i64(B, L)}
        self.P, self.cfg, self.with_user_proj = P, cfg, with_user_proj
        self.A = torch.empty(B, 2, 32, device=dev, dtype=torch.float32) if with_user_proj else None

        def body():
            probs, ctx = hotpath.forward(P, cfg, self.inp["i_node"], self.inp["neg_samples"], self.inp["seq_d1"],
                                         self.inp["seq_d2"], train=False, seed=0, dist=None, need_ctx=True)
            if with_user_proj:
                call("amid_catalogue_user_proj", _ptr(ctx.us[0]), _ptr(ctx.us[1]), B, _ptr(P["predictModule.fc.0.weight"]), 32,
                     _ptr(self.A), _stream())
            return probs, ctx

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                      # warm-up outside capture (lazy allocations, attributes)
            body()
            body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.probs, self.ctx = body()

    def run(self, batch: Dict[str, torch.Tensor]):
        for k, t in self.inp.items():
            t.copy_(batch[k].reshape(t.shape), non_blocking=True)
        self.graph.replay()
        return self.probs


@torch.no_grad()
def evaluate_full_catalogue_fast(P, cfg, cat: Catalogue, batches, dist=None, graph: bool = True) -> Dict[str, tuple]:
    """Same result as ``evaluate_full_catalogue`` without per-batch host synchronisation (BASELINE config 5 at scale):
    per batch only the (graph-replayed) eval forward and the user projection run, all user rows of this rank are then
    ranked against their pool by ONE k_rank_full launch per domain, and one device->host copy brings the counts back.
    Ties are resolved once at the end with the reference's numpy expression.  With ``dist`` every rank passes its own
    whole batches (the multi-interest module couples the users of a batch); the rank lists are combined with one
    padded tensor all-gather in rank order."""
    from . import hotpath
    batches = list(batches)
    if not batches:
        return {}
    dev = batches[0]["i_node"].device
    B, L = batches[0]["seq_d1"].shape
    nb = len(batches)
    A_all = torch.empty(nb * B, 2, 32, device=dev, dtype=torch.float32)
    dom_all = torch.empty(nb * B, device=dev, dtype=torch.int64)
    item_all = torch.empty(nb * B, device=dev, dtype=torch.int64)
    ov_all = torch.empty(nb * B, device=dev, dtype=torch.int64) if "overlap_label" in batches[0] else None
    gf = GraphedForward(P, cfg, B, L, 2, with_user_proj=True) if graph else None
    for k, b in enumerate(batches):
        sl = slice(k * B, (k + 1) * B)
        neg = b["neg_samples"][:, :1].contiguous() if "neg_samples" in b else b["i_node"].view(B, 1)
        if gf is not None:
            gf.run({"i_node": b["i_node"], "neg_samples": neg, "seq_d1": b["seq_d1"], "seq_d2": b["seq_d2"]})
            A_all[sl].copy_(gf.A)
        else:
            _, ctx = hotpath.forward(P, cfg, b["i_node"], neg, b["seq_d1"], b["seq_d2"], train=False, seed=0, dist=None,
                                     need_ctx=True)
            call("amid_catalogue_user_proj", _ptr(ctx.us[0]), _ptr(ctx.us[1]), B, _ptr(P["predictModule.fc.0.weight"]), 32,
                 _ptr(A_all[sl]), _stream())
        dom_all[sl].copy_(b["domain_id"])
        item_all[sl].copy_(b["i_node"])
        if ov_all is not None:
            ov_all[sl].copy_(b["overlap_label"])
    w2, b2 = P["predictModule.fc.2.weight"], P["predictModule.fc.2.bias"]
    fix32 = float(np.float32(FIX_VALUE))
    s = _stream()
    per_dom = {}
    for dom in (0, 1):
        rows = torch.nonzero(dom_all == dom).flatten().to(torch.int32)      # one host sync per domain per evaluation
        n = rows.numel()
        if n == 0:
            continue
        pos_idx = cat.index_of[dom][item_all].contiguous()
        lo, hi = cat.ranges[dom]
        counts = torch.empty(n, 4, device=dev, dtype=torch.int32)
        s_pos = torch.empty(n, device=dev, dtype=torch.float32)
        call("amid_catalogue_rank", _ptr(A_all), _ptr(rows), n, dom, _ptr(cat.Bc), lo, hi, _ptr(pos_idx), _ptr(w2), _ptr(b2),
             fix32, _ptr(counts), _ptr(s_pos), s)
        per_dom[dom] = (rows, pos_idx, counts, s_pos)
    acc = {}
    for dom, (rows, pos_idx, counts, s_pos) in per_dom.items():
        if bool((pos_idx[rows.long()] < 0).any()):
            raise IndexError(f"evaluate_full_catalogue_fast: a positive item of domain {dom + 1} is not in its pool")
        lo, hi = cat.ranges[dom]
        c = counts.cpu().numpy().astype(np.int64)
        ranks_nofix, ranks_fix = c[:, 0].copy(), c[:, 2].copy()
        tied = np.nonzero((c[:, 1] > 0) | (c[:, 3] > 0))[0]
        for t0 in range(0, len(tied), 256):               # resolve with the reference's own numpy expression (utils.py:297)
            tt = tied[t0:t0 + 256]
            sel = rows[torch.from_numpy(tt).to(dev)].contiguous()
            sc = torch.empty(len(tt), hi - lo, device=dev, dtype=torch.float32)
            sp = torch.empty(len(tt), device=dev, dtype=torch.float32)
            call("amid_catalogue_scores", _ptr(A_all), _ptr(sel), len(tt), dom, _ptr(cat.Bc), lo, hi, _ptr(pos_idx), _ptr(w2),
                 _ptr(b2), _ptr(sp), _ptr(sc), s)
            sc, sp = sc.cpu().numpy(), sp.cpu().numpy()
            pcol = (pos_idx[sel.long()] - lo).cpu().numpy()

            def resolve(k):
                others = np.delete(sc[k], pcol[k]) if 0 <= pcol[k] < hi - lo else sc[k]
                out = []
                for fixv, has_tie in ((0.0, c[tt[k], 1] > 0), (FIX_VALUE, c[tt[k], 3] > 0)):
                    if not has_tie:
                        out.append(None)
                        continue
                    row = np.concatenate((np.array([sp[k]], dtype=np.float32), others))
                    row[0] = row[0] - fixv
                    # (-row).argsort().argsort()[0] == position of index 0 in (-row).argsort(): same sort, same tie order
                    out.append(int(np.nonzero((-row).argsort() == 0)[0][0]))
                return out

            for k, (r_nofix, r_fix) in enumerate(_pool().map(resolve, range(len(tt)))):
                if r_nofix is not None:
                    ranks_nofix[tt[k]] = r_nofix
                if r_fix is not None:
                    ranks_fix[tt[k]] = r_fix
        name = "d1" if dom == 0 else "d2"
        acc[name] = ranks_fix
        if ov_all is not None:
            o = ov_all[rows.long()].cpu().numpy() != 0
            acc[name + "_ov"], acc[name + "_no"] = ranks_nofix[o], ranks_nofix[~o]
    res = {}
    for k in ("d1", "d2", "d1_ov", "d1_no", "d2_ov", "d2_no"):
        ranks = acc.get(k, np.zeros(0, dtype=np.int64))
        if dist is not None and dist.world > 1:
            ranks = _gather_ranks(ranks, dist, dev)
        if len(ranks):
            res[k] = metrics_from_ranks(ranks)
    return res


_POOL = None


def _pool():
    """Host threads for the numpy tie-breaking sorts (numpy releases the GIL inside argsort)."""
    global _POOL
    if _POOL is None:
        import os
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=max(1, min(16, (os.cpu_count() or 2) // 2)))
    return _POOL


def _gather_ranks(ranks: np.ndarray, dist, dev) -> np.ndarray:
    """Concatenate every rank's int64 list in rank order: one size all-gather + one padded tensor all-gather (NCCL)."""
    n = torch.tensor([len(ranks)], device=dev, dtype=torch.int64)
    sizes = torch.empty(dist.world, device=dev, dtype=torch.int64)
    dist.all_gather_into(sizes, n)
    sizes = sizes.cpu().numpy()
    cap = int(sizes.max()) if len(sizes) else 0
    if cap == 0:
        return ranks
    mine = torch.zeros(cap, device=dev, dtype=torch.int64)
    mine[:len(ranks)] = torch.from_numpy(np.ascontiguousarray(ranks)).to(dev)
    allr = torch.empty(dist.world * cap, device=dev, dtype=torch.int64)
    dist.all_gather_into(allr, mine)
    allr = allr.cpu().numpy().reshape(dist.world, cap)
    return np.concatenate([allr[r, :sizes[r]] for r in range(dist.world)])


# --------------------------------------------------------------------------------------------------
# Opt-in fast driver loop: test() of train_sr.py:31-128 without the per-field .cuda() calls,
# the D2H of every score matrix and the O(N^2) np.append bookkeeping -- same return tuple.
# --------------------------------------------------------------------------------------------------
_METRIC_ORDER_OVERLAP = ("d1_ov", "d1_no", "d2_ov", "d2_no", "d1", "d2")


@torch.no_grad()
def test(trainer, loader, overlap: bool = False):
    """Drop-in for ``test(model, args, valLoader)``: returns ``(loss, loss_cls, <7 metrics per list>...)`` in the
    reference's order -- lists (d1, d2) when ``overlap`` is false (train_sr.py:114-118), otherwise
    (d1_ov, d1_no, d2_ov, d2_no, d1, d2) (:119-128).  ``loader`` yields the reference's collated batches
    (dataset_seq.collate_fn_enhance); the last partial batch is expected to be dropped by the loader as in
    train_sr.py:455.  Under data parallelism each rank evaluates the whole batches of ITS loader on its own replica and
    gets the metrics of those batches (shard the eval set across ranks and combine the rank lists if needed)."""
    from . import hotpath
    was_training = trainer.model.training
    trainer.model.eval()
    trainer.flush()
    p1s, p2s, doms, ovs, losses = [], [], [], [], []
    for host in loader:
        b = trainer.to_device(host)
        probs = trainer.scores(b, local=True)
        B = probs.shape[2]
        l, _ = hotpath.loss_fwd_bwd(probs[:1].contiguous(), b["label"], b["domain_id"], None, 0, 0.0, B)
        losses.append(l[:1])
        p1s.append(probs[0, 0]); p2s.append(probs[0, 1]); doms.append(b["domain_id"])
        if overlap:
            ovs.append(b["overlap_label"])
    if was_training:
        trainer.model.train()
    if not p1s:
        raise ValueError("test(): empty loader")
    loss = float(np.mean(torch.cat(losses).cpu().numpy().astype(np.float64)))   # AverageMeter of loss.item()
    res = evaluate_lists(torch.cat(p1s), torch.cat(p2s), torch.cat(doms), torch.cat(ovs) if overlap else None)
    out = [loss, loss]
    for k in (_METRIC_ORDER_OVERLAP if overlap else ("d1", "d2")):
        out += list(res.get(k, (0.0,) * 7))
    return tuple(out)

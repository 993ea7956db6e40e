"""Fused training / evaluation engine around the drop-in ``SASRec``.

``Trainer.step`` is the whole reference training step (train_sr.py:201-215, or the two
phases of train_sr_dr.py:205-225 / 378-398) executed as one kernel sequence over the C
ABI: forward -> fused loss -> backward -> deterministic sort+segmented embedding-gradient
reduction -> Adam (one fused launch over the flat dense-parameter buffer, and a
row-sparse Adam on the table with exact dense-Adam semantics).
Under data parallelism (one process per GPU) the dense gradients are all-reduced in one
NCCL call and the locally pre-reduced table-gradient rows are all-gathered and reduced
again in fixed rank order.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _abi, hotpath
from ._abi import call
from .hotpath import D, _ptr, _stream

TABLE = "item_emb_layer.emb_item.weight"
BATCH_KEYS = ("i_node", "neg_samples", "seq_d1", "seq_d2", "domain_id", "label")


def pad_for_exchange(uid: torch.Tensor, ug: torch.Tensor, nu: torch.Tensor, V: int):
    """Fixed-size (row id, gradient row) list for the all-gather: slots past n_uniq get the
    sentinel id V-1 with a zero row, which a dense Adam would also step with zero gradient."""
    valid = torch.arange(uid.numel(), device=uid.device) < nu.to(torch.int64)
    uid = torch.where(valid, uid, torch.full_like(uid, V - 1))
    ug = torch.where(valid.unsqueeze(1), ug, torch.zeros_like(ug))
    return uid, ug


def owner_send_counts(uid: torch.Tensor, nu: torch.Tensor, Vs: int, G: int):
    """Routing plan of the sparse reduce-scatter (`Trainer._dense_table_step`).  ``uid[:nu]`` are this rank's touched row
    ids in ascending order (the output of the segmented reduction), rank o owns rows [o Vs, (o+1) Vs).  Returns the id list
    with its unused tail replaced by the sentinel G Vs (sorts behind every owner) and the number of rows destined to each
    owner -- contiguous runs of the list, because it is sorted.  Pure torch: runs on CPU tensors too (tests/test_dp_gloo.py)."""
    n = uid.numel()
    dev = uid.device
    valid = torch.arange(n, device=dev) < nu.to(torch.int64)
    ids = torch.where(valid, uid, torch.full_like(uid, G * Vs))
    bounds = torch.arange(1, G + 1, device=dev, dtype=torch.int64) * Vs
    ends = torch.searchsorted(ids, bounds)                 # ids < bound (negative = flagged ids: owner 0 skips them)
    send = torch.diff(ends, prepend=torch.zeros(1, device=dev, dtype=ends.dtype)).to(torch.int64)
    return ids, send


class _AdamState:
    def __init__(self, flat_numel: int, V: int, dev):
        z = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float32)
        self.m, self.v = z(flat_numel), z(flat_numel)
        self.tm, self.tv = z(V, D), z(V, D)
        self.last = torch.zeros(V, device=dev, dtype=torch.int32)
        self.step = 0


class Trainer:
    """Fused train step for an ``amid_b200.model_seq.SASRec`` on the current CUDA device."""

    def __init__(self, model, lr: float = 5e-4, lr2: float = 1.0, dr_e_w: float = 0.01, betas=(0.9, 0.999),
                 eps: float = 1e-8, dist: Optional[hotpath.DistCtx] = None, sparse_table: bool = True,
                 table_sync: str = "auto", rows_per_step_hint: Optional[int] = None):
        self.model, self.cfg, self.dist = model, model.cfg, dist
        self.lr, self.lr2, self.dr_e_w, self.betas, self.eps = lr, lr * lr2, dr_e_w, betas, eps   # train_sr_dr.py:668-669
        self.sparse_table = sparse_table
        P = model.param_dict()
        self.table = P[TABLE]
        if not self.table.is_cuda:
            raise _abi.AmidError("Trainer needs the model on a CUDA device (no CPU fallback)")
        dev = self.table.device
        # re-point every dense parameter at a view of one flat buffer (single Adam launch, single all-reduce)
        self.names = [n for n in P if n != TABLE]
        total = sum((P[n].numel() + 3) // 4 * 4 for n in self.names)
        self.flat_p = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(total, device=dev, dtype=torch.float32)
        self.G: Dict[str, torch.Tensor] = {}
        off = 0
        for n in self.names:
            k = P[n].numel()
            view = self.flat_p[off:off + k].view(P[n].shape)
            view.copy_(P[n].data)
            P[n].data = view
            self.G[n] = self.flat_g[off:off + k].view(P[n].shape)
            off += (k + 3) // 4 * 4
        self.P = {n: p.data for n, p in model.named_parameters()}
        self.V = self.table.shape[0]
        self.world = dist.world if dist is not None else 1
        # How replicas keep the replicated table identical under data parallelism:
        #  "sparse": all-gather the locally pre-reduced (row id, gradient row) lists, reduce again, row-sparse Adam
        #            (right when a step touches few distinct rows: real data, pad-heavy histories);
        #  "dense" : scatter into a dense [V,128] gradient, reduce-scatter it, dense Adam on this rank's row shard
        #            (Adam state sharded 1/N), all-gather the updated shard (right when most of the table is touched).
        if table_sync == "auto":
            rows = rows_per_step_hint if rows_per_step_hint is not None else 0
            table_sync = "dense" if (self.world > 1 and rows * self.world * 2 >= self.V) else "sparse"
        if table_sync == "sharded" and dist is None:
            raise _abi.AmidError("table_sync='sharded' needs a DistCtx (a process group, possibly of size 1)")
        self.table_sync = table_sync if (self.world > 1 or table_sync == "sharded") else "local"
        self.V_local = self.V
        self.sharded = None
        n_opt = 2 if self.cfg.isDR else 1                        # optimizer2 (train_sr_dr.py:669)
        if self.table_sync == "dense":
            self.Vs = (self.V + self.world - 1) // self.world     # rows per shard
            self.Vp = self.Vs * self.world
            padded = torch.zeros(self.Vp, D, device=dev, dtype=torch.float32)
            padded[:self.V].copy_(self.table.data)
            self.table_padded = padded
            self.table.data = padded[:self.V]                     # the parameter stays a [V,128] view
            self.shard_g = torch.empty(self.Vs, D, device=dev, dtype=torch.float32)
            self.opt = [_AdamState(total, 1, dev) for _ in range(n_opt)]
            for st in self.opt:
                st.tm = torch.zeros(self.Vs, D, device=dev, dtype=torch.float32)
                st.tv = torch.zeros(self.Vs, D, device=dev, dtype=torch.float32)
            self.sparse_table = False
        elif self.table_sync == "sharded":
            # large-vocabulary mode (BASELINE config 4): this rank keeps rows {rank, rank+G, ...} and their Adam
            # state only; every step fetches the rows it reads with one all-to-all lookup (amid_b200/sharded.py)
            from .sharded import ShardedTable
            self.sharded = ShardedTable.from_full(self.table.data, dist.rank, dist.world, dist.group)
            dist.sharded = self.sharded
            self.table.data = self.sharded.shard                   # the full table is released here
            self.V_local = self.sharded.Vs
            self.opt = [_AdamState(total, self.V_local, dev) for _ in range(n_opt)]
        else:
            self.opt = [_AdamState(total, self.V, dev) for _ in range(n_opt)]
        self.P = {n: p.data for n, p in model.named_parameters()}   # (re)bind after any re-pointing above
        self.table = self.P[TABLE]
        self.active_opt = 0
        self.last_losses = None
        self._seed = 0
        if self.world > 1:
            # replicas must start from rank 0's weights and share the dropout seed stream (the keep bits are indexed by
            # global sample, so every rank must hash with the same seed)
            tdist = dist.dist
            tdist.broadcast(self.flat_p, 0, group=dist.group)
            if self.table_sync != "sharded":
                tdist.broadcast(self.table_padded if self.table_sync == "dense" else self.table.data, 0, group=dist.group)
            sb = torch.tensor([model._seed_base], device=dev, dtype=torch.int64)
            tdist.broadcast(sb, 0, group=dist.group)
            model._seed_base = int(sb.item())

    # ------------------------------------------------------------------ helpers
    def _table_adam(self, st: _AdamState, uid, ug, nu, lr):
        call("amid_adam_rows_lazy", _ptr(self.table.data), _ptr(st.tm), _ptr(st.tv), _ptr(st.last), _ptr(uid), _ptr(ug),
             _ptr(nu), uid.numel(), st.step, lr, self.betas[0], self.betas[1], self.eps, _stream())

    def check_ids(self):
        """Raise IndexError if any kernel since the last check met an item id outside [0, V) (the reference raises
        at the embedding lookup; here the kernels skip the row, set a device flag, and the host polls it at flush /
        checkpoint / end of epoch -- one 4-byte read, no per-step synchronisation)."""
        code = _abi.lib().amid_gather_error_host_sync()
        if code:
            raise IndexError("amid_b200: an item id outside [0, item_length) reached the table kernels "
                             f"(device flag {code}); the offending rows were skipped")

    def flush(self):
        """Apply the pending zero-gradient Adam steps to every table row (exact dense semantics)."""
        self.check_ids()
        for i, st in enumerate(self.opt):
            if st.step > 0 and self.sparse_table:
                lr = self.lr if i == 0 else self.lr2
                call("amid_adam_rows_flush", _ptr(self.table.data), _ptr(st.tm), _ptr(st.tv), _ptr(st.last), self.V_local,
                     st.step, lr, self.betas[0], self.betas[1], self.eps, _stream())

    def to_device(self, host_batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """H2D of one batch (pinned host tensors recommended), ids as int64, labels fp32."""
        dev = self.table.device
        out = {}
        for k, v in host_batch.items():
            if k == "label":
                out[k] = v.to(dev, dtype=torch.float32, non_blocking=True)
            else:
                out[k] = v.to(dev, dtype=torch.int64, non_blocking=True)
        return out

    # ------------------------------------------------------------------ the step
    def step(self, batch: Dict[str, torch.Tensor], phase: int = 1) -> torch.Tensor:
        """One optimisation step.  phase 1: loss_cls (+ dr_e_w * loss_dr_e when isDR) with
        `optimizer`; phase 2 (isDR only): loss_dr_r with `optimizer2`.  Returns the device
        tensor [loss_cls, loss_dr_e, loss_dr_r] (local-batch share of the global means)."""
        cfg = self.cfg
        oi = 0 if phase == 1 else 1
        if phase == 2 and not cfg.isDR:
            raise _abi.AmidError("phase 2 needs isDR=True (train_sr_dr.py:381-384)")
        if oi != self.active_opt:
            self.flush()                       # the two Adams interleave on the same parameters
            self.active_opt = oi
        st = self.opt[oi]
        lr = self.lr if oi == 0 else self.lr2
        self._seed += 1
        seed = (self.model._seed_base * 1000003 + self._seed) & (2**63 - 1)
        train = self.model.training
        probs, ctx = hotpath.forward(self.P, cfg, batch["i_node"], batch["neg_samples"], batch["seq_d1"],
                                     batch["seq_d2"], train=train, seed=seed, dist=self.dist)
        mode = 0 if not cfg.isDR else (1 if phase == 1 else 2)
        B = probs.shape[2]
        losses, dprobs = hotpath.loss_fwd_bwd(probs, batch["label"], batch["domain_id"], batch.get("ob_label"), mode,
                                              self.dr_e_w, B * self.world)
        _, ids_all, rows_all = hotpath.backward(self.P, cfg, ctx, dprobs, G=self.G, dist=self.dist)
        if self.table_sync == "sharded":
            # ids_all are rows of the step table: the local reduction yields one gradient row per step-table row, in
            # bucket order; the owners reduce what all ranks send them and update their shard
            U = ctx.route.rows.shape[0]
            _, ug, _ = hotpath.segreduce(ids_all, rows_all, U)
            recv = self.sharded.push_grads(ctx.route, ug[:U])
            if self.world > 1:
                self.dist.all_reduce(self.flat_g)
                self.dist.all_reduce(losses)
            st.step += 1
            call("amid_adam_dense", _ptr(self.flat_p), _ptr(self.flat_g), _ptr(st.m), _ptr(st.v), self.flat_p.numel(),
                 st.step, lr, self.betas[0], self.betas[1], self.eps, _stream())
            if recv.shape[0]:
                uid, ug, nu = hotpath.segreduce(ctx.route.recv_local, recv, self.V_local)
                self._table_adam(st, uid, ug, nu, lr)
            self.last_losses = losses
            return losses
        uid, ug, nu = hotpath.segreduce(ids_all, rows_all, self.V)
        if self.world > 1:
            self.dist.all_reduce(self.flat_g)                      # dense grads: one NCCL call
            self.dist.all_reduce(losses)
            if self.table_sync == "dense":
                st.step += 1
                call("amid_adam_dense", _ptr(self.flat_p), _ptr(self.flat_g), _ptr(st.m), _ptr(st.v),
                     self.flat_p.numel(), st.step, lr, self.betas[0], self.betas[1], self.eps, _stream())
                self._dense_table_step(st, uid, ug, nu, lr)
                self.last_losses = losses
                return losses
            uid, ug, nu = self._exchange_table_grads(uid, ug, nu)
        st.step += 1
        call("amid_adam_dense", _ptr(self.flat_p), _ptr(self.flat_g), _ptr(st.m), _ptr(st.v), self.flat_p.numel(),
             st.step, lr, self.betas[0], self.betas[1], self.eps, _stream())
        if self.sparse_table:
            self._table_adam(st, uid, ug, nu, lr)
        else:                                                      # reference-equivalent dense pass over [V,128]
            dense = hotpath.dense_table_grad(uid, ug, nu, self.V)
            call("amid_adam_dense", _ptr(self.table.data), _ptr(dense), _ptr(st.tm), _ptr(st.tv), self.table.numel(),
                 st.step, lr, self.betas[0], self.betas[1], self.eps, _stream())
        self.last_losses = losses
        return losses

    def _dense_table_step(self, st, uid, ug, nu, lr):
        """Replicated table, most rows touched: SPARSE reduce-scatter of the table gradient -> dense Adam on this rank's row
        shard (Adam state sharded 1/N) -> all-gather of the updated rows.  Exactly torch.optim.Adam's dense semantics.

        The locally pre-reduced list (uid ascending, ug) is already bucketed by owner (rank r owns rows [r Vs, (r+1) Vs)), so
        each rank sends every owner only the rows it touched: an all-to-all of ~(N-1)/N x rows_touched x 512 B instead of a
        dense reduce-scatter of the whole 458 MB gradient.  The owner adds the received segments in rank order (unique ids
        inside a segment: no atomics), which fixes the summation order.  One [N,N] count exchange + host read per step."""
        import torch.distributed as tdist
        G, r, Vs = self.world, self.dist.rank, self.Vs
        dev = ug.device
        ids, send = owner_send_counts(uid, nu, Vs, G)
        counts = torch.empty(G * G, device=dev, dtype=torch.int64)
        tdist.all_gather_into_tensor(counts, send, group=self.dist.group)
        cm = counts.view(G, G).cpu()                                          # cm[s, o] = rows rank s sends to owner o
        in_splits = cm[r].tolist()
        out_splits = cm[:, r].tolist()
        n_in, n_out = int(sum(in_splits)), int(sum(out_splits))
        rid = torch.empty(n_out, device=dev, dtype=torch.int64)
        rrows = torch.empty(n_out, D, device=dev, dtype=torch.float32)
        tdist.all_to_all_single(rid, ids[:n_in].contiguous(), out_splits, in_splits, group=self.dist.group)
        tdist.all_to_all_single(rrows, ug[:n_in].contiguous(), out_splits, in_splits, group=self.dist.group)
        self.shard_g.zero_()
        off = 0
        for s_rank in range(G):                                               # fixed order -> deterministic sum
            c = out_splits[s_rank]
            if c:
                call("amid_embgrad_scatter_add", _ptr(rid[off:off + c]), _ptr(rrows[off:off + c]), c, r * Vs, _ptr(self.shard_g), Vs,
                     _stream())
            off += c
        shard_p = self.table_padded[r * Vs:(r + 1) * Vs]                      # Adam in place on the owned rows
        call("amid_adam_dense", _ptr(shard_p), _ptr(self.shard_g), _ptr(st.tm), _ptr(st.tv), shard_p.numel(),
             st.step, lr, self.betas[0], self.betas[1], self.eps, _stream())
        tdist.all_gather_into_tensor(self.table_padded, shard_p, group=self.dist.group)    # in place (NCCL: sendbuff = recvbuff + rank * count)

    def _exchange_table_grads(self, uid, ug, nu):
        """Replicated table under DP: all-gather every rank's pre-reduced (row id, gradient row) list and reduce again
        in rank order, so every replica applies the identical update.  The lists are cut to the largest per-rank count
        (one small all-gather + host read per step) -- on real data a step touches a few per cent of its positions'
        worth of distinct rows (the pad row dominates), so the exchange is megabytes, not the 210 MB of a full-length
        list; unused slots carry the sentinel id V-1 with a zero row."""
        n = uid.numel()
        dev = ug.device
        nu_all = torch.empty(self.world, device=dev, dtype=torch.int32)
        self.dist.all_gather_into(nu_all, nu.reshape(1).to(torch.int32))
        cap = min(n, max(256, (int(nu_all.max()) + 255) // 256 * 256))
        uid, ug = pad_for_exchange(uid[:cap], ug[:cap], nu, self.V)
        all_ids = torch.empty(cap * self.world, device=dev, dtype=torch.int64)
        all_rows = torch.empty(cap * self.world, D, device=dev, dtype=torch.float32)
        self.dist.all_gather_into(all_ids, uid.contiguous())
        self.dist.all_gather_into(all_rows, ug.contiguous())
        return hotpath.segreduce(all_ids, all_rows, self.V)

    # ------------------------------------------------------------------ evaluation forward
    @torch.no_grad()
    def scores(self, batch: Dict[str, torch.Tensor], local: bool = False) -> torch.Tensor:
        """Eval-mode probabilities [n_heads, 2, B, C] for a batch (model.eval() semantics).  By default the batch is
        this rank's slice of a global batch (the multi-interest module couples the whole batch).  ``local=True`` scores a
        WHOLE batch on this rank alone -- how evaluation shards work across ranks (every rank takes different whole
        batches); it needs the replicated table."""
        dist = self.dist
        if local and self.world > 1:
            if self.sharded is not None:
                raise _abi.AmidError("local scoring needs the replicated table; with table_sync='sharded' every rank "
                                     "has to take part in the lookup")
            dist = None
        probs, _ = hotpath.forward(self.P, self.cfg, batch["i_node"], batch["neg_samples"], batch["seq_d1"],
                                   batch["seq_d2"], train=False, seed=0, dist=dist, need_ctx=False)
        return probs

    # ------------------------------------------------------------------ full-catalogue evaluation (config 5)
    def catalogue(self, pool_d1: torch.Tensor, pool_d2: torch.Tensor):
        """Item halves of the scorer for both domain pools; pending lazy-Adam rows are flushed first."""
        from . import evaluate
        if self.sharded is not None:
            raise _abi.AmidError("full-catalogue evaluation reads the replicated table; with table_sync='sharded' rebuild "
                                 "one with full_table() and evaluate through a replicated-table Trainer")
        self.flush()
        return evaluate.Catalogue(self.P, self.cfg, pool_d1, pool_d2)

    def evaluate_full_catalogue(self, cat, batches, fast: bool = True):
        """HR/NDCG/MRR of every user against the whole pool of its target domain.  Under data parallelism each
        rank passes its own contiguous block of whole eval batches; the rank lists are gathered in rank order.
        ``fast`` = graph-replayed forwards, one rank launch per domain, one read-back (same result)."""
        from . import evaluate
        fn = evaluate.evaluate_full_catalogue_fast if fast else evaluate.evaluate_full_catalogue
        return fn(self.P, self.cfg, cat, list(batches), self.dist)

    def full_table(self) -> torch.Tensor:
        """The whole [V,128] item table on every rank (checkpoint / state_dict); pending lazy-Adam rows are flushed."""
        self.flush()
        return self.sharded.full_table() if self.sharded is not None else self.table.data[:self.V]

    # ------------------------------------------------------------------ checkpoint / resume
    def checkpoint(self) -> Dict[str, object]:
        """Everything needed to resume bit-exactly: parameters by their reference state-dict names (pending lazy-Adam
        rows flushed first; a row-sharded table is saved as this rank's shard), both optimizers' moments and step
        counters, and the dropout seed state.  The reference's own checkpointing is commented out
        (train_sr.py:181-186, 327-332); the only contract inherited from it is the state-dict naming."""
        self.flush()
        ck = {"format": 1, "table_sync": self.table_sync, "world": self.world,
              "rank": self.dist.rank if self.dist is not None else 0,
              "params": {n: p.detach().clone() for n, p in self.P.items()},
              "opt": [{"m": st.m.clone(), "v": st.v.clone(), "tm": st.tm.clone(), "tv": st.tv.clone(),
                       "last": st.last.clone(), "step": st.step} for st in self.opt],
              "active_opt": self.active_opt, "seed": self._seed, "seed_base": self.model._seed_base,
              "model_step": self.model._step}
        return ck

    def load_checkpoint(self, ck: Dict[str, object]) -> None:
        if ck.get("format") != 1:
            raise _abi.AmidError("unknown checkpoint format")
        if ck["table_sync"] != self.table_sync or ck["world"] != self.world:
            raise _abi.AmidError(f"checkpoint was written with table_sync={ck['table_sync']} on {ck['world']} rank(s); "
                                 f"this trainer runs table_sync={self.table_sync} on {self.world}")
        my_rank = self.dist.rank if self.dist is not None else 0
        if self.table_sync in ("sharded", "dense") and int(ck.get("rank", 0)) != my_rank:
            raise _abi.AmidError(f"checkpoint was written by rank {ck.get('rank')}: with table_sync={self.table_sync} the table "
                                 f"shard / Adam moments are per rank (one file per rank); this is rank {my_rank}")
        if len(ck["opt"]) != len(self.opt):
            raise _abi.AmidError("checkpoint and model disagree on isDR (number of optimizers)")
        for n, p in self.P.items():
            src = ck["params"][n]
            if tuple(src.shape) != tuple(p.shape):
                raise _abi.AmidError(f"checkpoint tensor {n} has shape {tuple(src.shape)}, expected {tuple(p.shape)}")
            p.copy_(src)
        for st, s in zip(self.opt, ck["opt"]):
            st.m.copy_(s["m"]); st.v.copy_(s["v"]); st.tm.copy_(s["tm"]); st.tv.copy_(s["tv"]); st.last.copy_(s["last"])
            st.step = int(s["step"])
        self.active_opt = int(ck["active_opt"])
        self._seed = int(ck["seed"])
        self.model._seed_base = int(ck["seed_base"])
        self.model._step = int(ck.get("model_step", self.model._step))

    # ------------------------------------------------------------------ opt-in fast epoch loop
    def train_epoch(self, loader, phase: int = 1, log_every: int = 0, log=print) -> float:
        """The inner loop of train() (train_sr.py:189-219 / train_sr_dr.py:189-229, 362-402) over the reference's
        collated batches: one fused step per batch, no per-step host synchronisation (the running loss is read back
        once at the end, or every ``log_every`` steps).  Returns the epoch's mean loss as the reference's AverageMeter
        would report it (mean of the per-batch loss of the phase)."""
        self.model.train()
        acc = None
        n = 0
        col = 0 if (phase == 1 and not self.cfg.isDR) else None
        for i, host in enumerate(loader):
            losses = self.step(self.to_device(host), phase=phase)
            if col is not None:
                cur = losses[col]
            elif phase == 1:
                cur = losses[0] + self.dr_e_w * losses[1]            # train_sr_dr.py:221
            else:
                cur = losses[2]                                       # :394
            acc = cur.double() if acc is None else acc + cur.double()
            n += 1
            if log_every and i % log_every == 0:
                log(f"train total loss:{float(acc) / n}")
        if n == 0:
            raise ValueError("train_epoch(): empty loader")
        self.check_ids()
        return float(acc) / n

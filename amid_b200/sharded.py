"""Row-sharded item table for the large-vocabulary configuration (BASELINE config 4, SURVEY.md 8e "Embedding
tables": ``owner = row mod G``; forward all-to-all(ids) -> all-to-all(rows), backward all-to-all(grad rows) to the
owners, owner does the segmented reduce + row-Adam).

Rank r of G holds rows {r, r+G, r+2G, ...} of ``emb_item.weight`` (model_seq.py:25) as a contiguous
[ceil(V/G), 128] shard plus the Adam state of those rows; no rank ever holds the whole table.

One lookup per step serves every table read of the step (candidates + both histories):

  1. the step's ids are de-duplicated (the left-pad row is ~75 % of all positions on real data);
  2. unique ids are bucketed by owner (stable, so the order inside a bucket is ascending id);
  3. NCCL all-to-all of the bucket sizes, then of the owner-local row indices;
  4. each owner gathers the requested rows from its shard (csrc/gather.cu, 128-bit row loads);
  5. NCCL all-to-all of the rows back.  The requester now holds a compact [U,128] "step table" and every position
     of the batch is re-labelled with its row in that table, so all downstream kernels (fused gather + positional
     add + mask, the backward's segmented reduction) run unchanged on ``(step table, virtual ids)``.

Backward: the segmented reduction over virtual ids yields one gradient row per step-table row, already in bucket
order; one all-to-all with the forward's split sizes delivers them to the owners, which reduce the rows of all ranks
again (fixed rank order => deterministic) and apply the lazy row-Adam to their shard.

The bucket sizes have to reach the host to size the all-to-all (one device->host read per lookup); this mode trades
that for never replicating a 10 GB table.  The routing is plain index arithmetic and is covered on CPU by
tests/test_dp_gloo.py with an injected gather; on a GPU the gather is the CUDA kernel and nothing else.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional

import torch

from . import _abi

D = 128


@dataclass
class Route:
    rows: torch.Tensor            # [U,128] step table: the unique rows this rank's batch reads, bucket order
    virtual_ids: torch.Tensor     # [R] int64: row of `rows` for every requested position
    send_splits: List[int]        # unique ids sent to each owner (forward) == gradient rows sent back (backward)
    recv_splits: List[int]        # requests received from each rank
    recv_local: torch.Tensor      # [sum(recv_splits)] int64 owner-local row indices requested from this rank


def _cuda_gather(shard: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """rows = shard[idx] through the C ABI (amid_emb_gather_fwd); refuses anything but CUDA tensors."""
    if not shard.is_cuda:
        raise _abi.AmidError("sharded table lookup needs CUDA tensors (no CPU fallback)")
    from .hotpath import _ptr, _stream
    out = torch.empty(idx.numel(), D, device=shard.device, dtype=torch.float32)
    if idx.numel():
        _abi.call("amid_emb_gather_fwd", _ptr(shard), shard.shape[0], _ptr(idx), idx.numel(), _ptr(out), _stream())
    return out


class ShardedTable:
    """This rank's rows of the item table and the exchange plan of one step."""

    def __init__(self, shard: torch.Tensor, V: int, rank: int, world: int, group=None,
                 gather: Optional[Callable[[torch.Tensor, torch.Tensor], torch.Tensor]] = None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world, self.V = rank, world, V
        self.Vs = (V + world - 1) // world
        if tuple(shard.shape) != (self.Vs, D):
            raise ValueError(f"shard must be [{self.Vs}, {D}], got {tuple(shard.shape)}")
        self.shard = shard
        self._gather = gather or _cuda_gather

    # -------------------------------------------------------------- construction helpers
    @staticmethod
    def rows_of(rank: int, world: int, V: int) -> torch.Tensor:
        return torch.arange(rank, V, world)

    @classmethod
    def from_full(cls, table: torch.Tensor, rank: int, world: int, group=None, gather=None) -> "ShardedTable":
        """Cut this rank's shard out of a full [V,128] table (tests, small vocabularies, checkpoint load)."""
        V = table.shape[0]
        Vs = (V + world - 1) // world
        shard = torch.zeros(Vs, D, device=table.device, dtype=table.dtype)
        mine = table[rank::world]
        shard[:mine.shape[0]].copy_(mine)
        return cls(shard, V, rank, world, group, gather)

    def full_table(self) -> torch.Tensor:
        """Reassemble the [V,128] table on every rank (state_dict / checkpoint; not on the step path)."""
        parts = [torch.empty_like(self.shard) for _ in range(self.world)]
        self.dist.all_gather(parts, self.shard.contiguous(), group=self.group)
        return torch.stack(parts, 1).reshape(self.Vs * self.world, D)[:self.V].contiguous()

    # -------------------------------------------------------------- the step's lookup
    def _plan(self, ids: torch.Tensor):
        """(send_local [U], virtual_ids [R], send_splits, recv_splits): the step's unique rows in bucket order, the row of
        the step table for every position, and the all-to-all split sizes.  On CUDA tensors this is ONE call into the
        library (amid_shard_plan: sort + run-length encode + scan + emit kernels) and one device->host read of
        2G + 2 integers; the torch expression below only serves the CPU routing test (tests/test_dp_gloo.py), which
        injects its own gather."""
        G = self.world
        if ids.is_cuda:
            from .hotpath import _ptr, _stream
            n = ids.numel()
            dev = ids.device
            uniq_local = torch.empty(n, device=dev, dtype=torch.int64)
            virt = torch.empty(n, device=dev, dtype=torch.int64)
            flags = torch.empty(2, device=dev, dtype=torch.int32)
            send_counts = torch.empty(G, device=dev, dtype=torch.int64)
            wsb = _abi.lib().amid_shard_plan_workspace_bytes(n)
            ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
            _abi.call("amid_shard_plan", _ptr(ids), n, self.V, G, _ptr(uniq_local), _ptr(virt), _ptr(flags), _ptr(send_counts),
                      _ptr(ws), wsb, _stream())
            recv_counts = torch.empty_like(send_counts)
            self.dist.all_to_all_single(recv_counts, send_counts, group=self.group)
            host = torch.cat((send_counts, recv_counts, flags.to(torch.int64))).tolist()   # the one host read of a lookup
            if host[-1]:
                raise IndexError("sharded lookup: item id out of range")
            send_splits, recv_splits = host[:G], host[G:2 * G]
            return uniq_local[:sum(send_splits)], virt, send_splits, recv_splits
        oob = ((ids < 0) | (ids >= self.V)).any().to(torch.int64).reshape(1)
        uniq, inv = torch.unique(ids, return_inverse=True)                 # sorted ascending
        dest = uniq % G
        order = torch.argsort(dest, stable=True)                           # bucket by owner, ascending id inside
        send_local = (uniq // G)[order].contiguous()
        send_counts = torch.bincount(dest, minlength=G)
        recv_counts = torch.empty_like(send_counts)
        self.dist.all_to_all_single(recv_counts, send_counts, group=self.group)
        host = torch.cat((send_counts, recv_counts, oob)).tolist()
        if host[-1]:
            raise IndexError("sharded lookup: item id out of range")
        slot_of = torch.empty_like(order)                                  # unique index -> row of the step table
        slot_of[order] = torch.arange(order.numel(), device=ids.device)
        return send_local, slot_of[inv].contiguous(), host[:G], host[G:2 * G]

    def lookup(self, ids: torch.Tensor) -> Route:
        if ids.dtype != torch.int64 or ids.dim() != 1:
            raise ValueError("ids must be a flat int64 tensor")
        send_local, virtual_ids, send_splits, recv_splits = self._plan(ids.contiguous())
        recv_local = torch.empty(sum(recv_splits), device=ids.device, dtype=torch.int64)
        self.dist.all_to_all_single(recv_local, send_local.contiguous(), recv_splits, send_splits, group=self.group)
        out_rows = self._gather(self.shard, recv_local)                    # owner side: csrc/gather.cu
        rows = torch.empty(sum(send_splits), D, device=ids.device, dtype=torch.float32)
        self.dist.all_to_all_single(rows, out_rows, send_splits, recv_splits, group=self.group)
        return Route(rows, virtual_ids, send_splits, recv_splits, recv_local)

    def push_grads(self, route: Route, grad_rows: torch.Tensor) -> torch.Tensor:
        """grad_rows [U,128], one per step-table row (bucket order) -> the gradient rows of every rank's requests to
        this owner, aligned with ``route.recv_local`` (rank order)."""
        if grad_rows.shape[0] != route.rows.shape[0]:
            raise ValueError("one gradient row per step-table row expected")
        recv = torch.empty(route.recv_local.numel(), D, device=grad_rows.device, dtype=torch.float32)
        self.dist.all_to_all_single(recv, grad_rows.contiguous(), route.recv_splits, route.send_splits, group=self.group)
        return recv

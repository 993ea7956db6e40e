"""world_size-2 gloo tests (CPU) of the data-parallel decomposition used by amid_b200.hotpath /
engine: batch-sharded MIM exchange and the table-gradient exchange.
The per-rank arithmetic is done with the oracle (checker only); what is under test is the
host-side sharding algebra and the DistCtx collectives."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from common import make_params
from oracle import amid_oracle as O

D = 128


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from amid_b200.engine import pad_for_exchange
        from amid_b200.hotpath import DistCtx
        ctx = DistCtx()
        torch.manual_seed(0)                       # same "global batch" on every rank
        Bg, n, ts = 8, 5, 0.15
        Bl, j0 = Bg // world, rank * (Bg // world)
        P = make_params(3, 4, D, n, 32, Bg)
        e1, e2 = torch.randn(Bg, n, D) * 0.3, torch.randn(Bg, n, D) * 0.3
        e1[2] *= 3; e2[2] *= 3; e1[6] *= 3; e2[6] = e1[6].clone()          # gates fire on both shards
        w_nn, b_nn = P["itc_d1.trans_nn.weight"], P["itc_d1.trans_nn.bias"]
        w_bs, b_bs = P["itc_d1.trans_bs.weight"], P["itc_d1.trans_bs.bias"]
        # ---- global (single-process) answer
        col = {}
        ref = O.mim_closed(e1, e2, w_nn, b_nn, w_bs, b_bs, ts, collect=col)
        assert 0 < col["g"].sum() < Bg
        # ---- sharded: all-gather m, global gate, local partial aggregate, all-reduce
        m_loc = O.mim_scores(e1[j0:j0 + Bl], e2[j0:j0 + Bl])
        m_glob = torch.empty(Bg)
        ctx.all_gather_into(m_glob, m_loc)
        assert torch.equal(m_glob, col["m"])
        g = (torch.softmax(m_glob, 0) > ts).float()
        coef = w_bs.reshape(-1) * g
        S = torch.einsum("j,jnd->nd", coef[j0:j0 + Bl], e2[j0:j0 + Bl])
        ctx.all_reduce(S)
        E = S @ w_nn.T + w_bs.sum() * b_nn + b_bs
        np.testing.assert_allclose(E.numpy(), col["E"].numpy(), rtol=0, atol=2e-6)
        np.testing.assert_allclose(E.numpy(), ref[0, n:].numpy(), rtol=0, atol=2e-6)
        # ---- backward of the shared half: dE = all-reduce of the local column sums; w_bs slices are disjoint
        du = torch.randn(Bg, D)
        dcol = du[j0:j0 + Bl].sum(0) / (2 * n)
        ctx.all_reduce(dcol)
        np.testing.assert_allclose(dcol.numpy(), (du.sum(0) / (2 * n)).numpy(), rtol=1e-5, atol=1e-6)
        dE = dcol.expand(n, D)
        dS = dE @ w_nn
        dw_loc = g[j0:j0 + Bl] * torch.einsum("nd,jnd->j", dS, e2[j0:j0 + Bl]) + (dE.sum(0) * b_nn).sum()
        gw = torch.zeros(Bg)
        gw[j0:j0 + Bl] = dw_loc
        ctx.all_reduce(gw)
        wv = w_bs.clone().requires_grad_(True)
        out = O.mim_closed(e1, e2, w_nn, b_nn, wv, b_bs, ts)
        (out[:, n:] * (du / (2 * n)).unsqueeze(1)).sum().backward()
        np.testing.assert_allclose(gw.numpy(), wv.grad.reshape(-1).numpy(), rtol=1e-4, atol=1e-5)
        # ---- table-gradient exchange: padded all-gather + second reduction == global reduction
        V = 50
        gen = torch.Generator().manual_seed(5)
        ids = torch.randint(0, V - 1, (world, 40), generator=gen)
        rows = torch.randn(world, 40, D, generator=gen)
        uid_l, inv = torch.unique(ids[rank], return_inverse=True)
        ug_l = torch.zeros(len(uid_l), D).index_add_(0, inv, rows[rank])
        k = len(uid_l)
        uid_p = torch.cat((uid_l, torch.full((40 - k,), 12345)))             # garbage past n_uniq
        ug_p = torch.cat((ug_l, torch.full((40 - k, D), float("nan"))))
        uid_x, ug_x = pad_for_exchange(uid_p, ug_p, torch.tensor([k], dtype=torch.int32), V)
        assert (uid_x[k:] == V - 1).all() and (ug_x[k:] == 0).all()
        all_ids, all_rows = torch.empty(40 * world, dtype=torch.int64), torch.empty(40 * world, D)
        ctx.all_gather_into(all_ids, uid_x)
        ctx.all_gather_into(all_rows, ug_x)
        dense = torch.zeros(V, D).index_add_(0, all_ids, all_rows)
        want = torch.zeros(V, D).index_add_(0, ids.reshape(-1), rows.reshape(-1, D))
        np.testing.assert_allclose(dense.numpy(), want.numpy(), rtol=1e-5, atol=1e-5)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_dp_decomposition_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}: {msg}"


# ---------------------------------------------------------------------------------------------------------------
# Row-sharded item table (BASELINE config 4): the all-to-all routing of amid_b200/sharded.py on two gloo ranks.
# The owner-side row gather is injected as a torch index (test scaffolding for the HOST logic only; on a GPU the
# class uses the CUDA gather kernel and refuses CPU tensors).
# ---------------------------------------------------------------------------------------------------------------
def _sharded_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from amid_b200 import _abi
        from amid_b200.sharded import ShardedTable
        V = 101                                             # not a multiple of the world size
        gen = torch.Generator().manual_seed(17)
        table = torch.randn(V, D, generator=gen)            # same on every rank
        sh = ShardedTable.from_full(table, rank, world, gather=lambda shard, idx: shard[idx])
        assert sh.shard.shape == ((V + world - 1) // world, D)
        assert torch.equal(sh.shard[:len(table[rank::world])], table[rank::world])
        # the default gather refuses CPU tensors: no silent fallback
        try:
            ShardedTable.from_full(table, rank, world).lookup(torch.zeros(3, dtype=torch.int64))
            raise AssertionError("CPU lookup must be refused")
        except _abi.AmidError:
            pass
        dist.barrier()
        # each rank asks for its own ragged, duplicate-heavy id list (pad-row style hot id 7)
        g2 = torch.Generator().manual_seed(100 + rank)
        ids = torch.cat((torch.randint(0, V, (37 + 5 * rank,), generator=g2), torch.full((20,), 7)))
        route = sh.lookup(ids)
        assert torch.equal(route.rows[route.virtual_ids], table[ids])               # every position gets its row
        U = len(torch.unique(ids))
        assert route.rows.shape == (U, D) and sum(route.send_splits) == U
        assert int(route.virtual_ids.max()) == U - 1
        # requests this owner received are its own rows, in rank order
        owned_global = route.recv_local * world + rank
        assert bool((owned_global < V).all())
        # backward: one gradient row per step-table row -> owners; scatter-add must equal the global dense gradient
        pos_grads = torch.randn(len(ids), D, generator=g2)
        step_grads = torch.zeros(U, D).index_add_(0, route.virtual_ids, pos_grads)
        recv = sh.push_grads(route, step_grads)
        assert recv.shape == (route.recv_local.numel(), D)
        mine = torch.zeros(sh.Vs, D).index_add_(0, route.recv_local, recv)           # owner-side reduction
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        got = torch.stack(parts, 1).reshape(sh.Vs * world, D)[:V]
        # reference: all ranks' position gradients scattered into a dense [V,128]
        all_ids, all_g = [None] * world, [None] * world
        dist.all_gather_object(all_ids, ids)
        dist.all_gather_object(all_g, pos_grads)
        want = torch.zeros(V, D)
        for i, gr in zip(all_ids, all_g):
            want.index_add_(0, i, gr)
        np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-5, atol=1e-5)
        # checkpoint view
        assert torch.equal(sh.full_table(), table)
        # empty request list on one rank must not dead-lock the collectives
        r2 = sh.lookup(ids[:0] if rank == 1 else ids[:5])
        assert r2.rows.shape[0] == (0 if rank == 1 else len(torch.unique(ids[:5])))
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_sharded_table_routing_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}: {msg}"


def _rs_worker(rank, world, port, q):
    """Sparse reduce-scatter of the table gradient (engine.Trainer._dense_table_step): routing plan + all-to-all + ordered
    owner-side accumulation against the dense sum of every rank's gradient."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from amid_b200.engine import owner_send_counts
        V, Dd = 37, 4
        Vs = (V + world - 1) // world
        g = torch.Generator().manual_seed(100 + rank)
        touched = torch.randperm(V, generator=g)[:11 + 3 * rank].sort().values          # this rank's unique rows, ascending
        nu = torch.tensor([touched.numel()], dtype=torch.int32)
        cap = 20
        uid = torch.full((cap,), 12345, dtype=torch.int64)                                # garbage behind n_uniq
        uid[:touched.numel()] = touched
        ug = torch.randn(cap, Dd, generator=g)
        ids, send = owner_send_counts(uid, nu, Vs, world)
        assert int(send.sum()) == touched.numel()
        for o in range(world):
            assert int(send[o]) == int(((touched >= o * Vs) & (touched < (o + 1) * Vs)).sum())
        counts = [torch.empty(world, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(counts, send)
        cm = torch.stack(counts)                                                          # cm[s, o]
        in_splits, out_splits = cm[rank].tolist(), cm[:, rank].tolist()
        n_in, n_out = sum(in_splits), sum(out_splits)
        rid, rrows = torch.empty(n_out, dtype=torch.int64), torch.empty(n_out, Dd)
        dist.all_to_all_single(rid, ids[:n_in].contiguous(), out_splits, in_splits)
        dist.all_to_all_single(rrows, ug[:n_in].contiguous(), out_splits, in_splits)
        shard = torch.zeros(Vs, Dd)
        off = 0
        for s_rank in range(world):                                                       # rank order, as amid_embgrad_scatter_add
            c = out_splits[s_rank]
            shard.index_add_(0, rid[off:off + c] - rank * Vs, rrows[off:off + c])
            off += c
        # dense reference: every rank's dense gradient, summed
        dense = torch.zeros(world * Vs, Dd)
        dense[touched] = ug[:touched.numel()]
        dist.all_reduce(dense)
        np.testing.assert_allclose(shard.numpy(), dense[rank * Vs:(rank + 1) * Vs].numpy(), rtol=0, atol=1e-6)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_sparse_reduce_scatter_routing_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rs_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res

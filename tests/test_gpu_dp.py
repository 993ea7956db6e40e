"""Multi-GPU data-parallel equivalence (needs >= 2 GPUs on the box; skipped otherwise)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(nproc, table_sync, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT,
                       env={**os.environ, "AMID_TABLE_SYNC": table_sync})
    assert "DP_CHECK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("table_sync", ["sparse", "dense", "sharded"])
def test_dp2_matches_single_gpu(table_sync):
    _run(2, table_sync, 29533)


def test_sharded_table_world1_matches_replicated():
    """The row-sharded table path (all-to-all lookup, step table, owner-side reduce + lazy Adam) on a one-rank
    process group: same steps as the plain single-GPU trainer."""
    _run(1, "sharded", 29535)


@pytest.mark.parametrize("n,V,G,pad_frac", [(1, 10, 1, 0.0), (1000, 97, 4, 0.0), (50000, 20_000_002, 8, 0.0), (30000, 1000, 2, 0.75),
                                            (4099, 5, 3, 0.0)])
def test_shard_plan_kernel_matches_torch_expression(n, V, G, pad_frac):
    """amid_shard_plan (sort + run-length encode + scan + emit) against the torch.unique / argsort / bincount plan of
    round 1: unique owner-local rows in bucket order, the step-table row of every position, the per-owner counts."""
    from amid_b200 import hotpath as hp
    from amid_b200._abi import call, lib
    g = torch.Generator().manual_seed(n + G)
    ids = torch.randint(0, V, (n,), generator=g)
    if pad_frac:
        ids[torch.rand(n, generator=g) < pad_frac] = V // 2
    ids = ids.cuda()
    uniq_local = torch.full((n,), -1, device="cuda", dtype=torch.int64)
    virt = torch.full((n,), -1, device="cuda", dtype=torch.int64)
    flags = torch.empty(2, device="cuda", dtype=torch.int32)
    counts = torch.empty(G, device="cuda", dtype=torch.int64)
    wsb = lib().amid_shard_plan_workspace_bytes(n)
    ws = torch.empty(wsb, device="cuda", dtype=torch.uint8)
    call("amid_shard_plan", hp._ptr(ids), n, V, G, hp._ptr(uniq_local), hp._ptr(virt), hp._ptr(flags), hp._ptr(counts),
         hp._ptr(ws), wsb, hp._stream())
    torch.cuda.synchronize()
    uniq, inv = torch.unique(ids, return_inverse=True)
    dest = uniq % G
    order = torch.argsort(dest, stable=True)
    want_local = (uniq // G)[order]
    slot_of = torch.empty_like(order)
    slot_of[order] = torch.arange(order.numel(), device="cuda")
    U = uniq.numel()
    assert flags.tolist() == [U, 0]
    assert torch.equal(uniq_local[:U], want_local)
    assert torch.equal(virt, slot_of[inv])
    assert torch.equal(counts, torch.bincount(dest, minlength=G))
    # an out-of-range id raises the flag and never aliases a valid row
    bad = ids.clone()
    bad[0] = V
    call("amid_shard_plan", hp._ptr(bad), n, V, G, hp._ptr(uniq_local), hp._ptr(virt), hp._ptr(flags), hp._ptr(counts),
         hp._ptr(ws), wsb, hp._stream())
    assert flags.tolist()[1] == 1

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

"""Multi-GPU data-parallel equivalence (needs >= 2 GPUs on the box; skipped otherwise)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("table_sync", ["sparse", "dense"])
def test_dp2_matches_single_gpu(table_sync):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "dp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT,
                       env={**os.environ, "AMID_TABLE_SYNC": table_sync})
    assert "DP_CHECK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]

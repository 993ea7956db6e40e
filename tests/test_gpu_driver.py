"""The reference's unmodified train_sr_dr.py (run.sh's driver) for one epoch per seed with amid_b200 as the drop-in
model_seq: its own DataLoaders (8 workers), samplers, two-phase DR training with two torch.optim.Adam instances and its
test() loop all run as shipped; the log must report finite losses and metrics for both domains."""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "train_sr_dr.py")), reason="oracle/_ref not built")
def test_unmodified_train_sr_dr_runs_with_the_dropin():
    cmd = [sys.executable, os.path.join(ROOT, "tools", "run_reference_driver.py"), "train_sr_dr.py", "--epoch", "1", "--model", "sasrec",
           "--isItC", "True", "--ts2", "0.4", "-ds", "amazon", "-dm", "cloth_sport", "--overlap_ratio", "0.75", "--neg_nums", "199",
           "--lr2", "0.01", "--dr_e_w", "0.01", "--overlap", "True"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=ROOT)
    out = r.stdout + r.stderr
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "r2_unmodified_driver.log"), "w") as fh:
        fh.write(out)
    assert r.returncode == 0, out[-4000:]
    assert "libamid_b200" not in out or "missing" not in out
    losses = [float(x) for x in re.findall(r"loss[^\n]*?([0-9]+\.[0-9]+)", out)]
    assert losses and all(0.0 < v < 50.0 for v in losses[:200]), losses[:10]
    assert re.search(r"(?i)hit|ndcg|mrr", out), out[-2000:]

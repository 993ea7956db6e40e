import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from amid_b200 import evaluate, hotpath, _abi
from amid_b200.engine import Trainer
from amid_b200.model_seq import SASRec
D, HID, V = 128, 32, 894820
B, L = 256, 20
n1, n2 = 16084, 12153
m = SASRec(0, D, V, D, L, HID, B, False, True, 0.5, 0.4, isDR=False).cuda().eval()
m.cfg.precision = "bf16"
tr = Trainer(m)
rng = np.random.default_rng(9)
perm = torch.from_numpy(rng.permutation(V)[:n1 + n2])
pool1, pool2 = perm[:n1], perm[n1:]
cat = tr.catalogue(pool1, pool2)
b = bench.synth_batch(rng, B, L, 2, V)
dom = b["domain_id"]
b["i_node"] = torch.where(dom == 0, pool1[torch.from_numpy(rng.integers(0, n1, B))], pool2[torch.from_numpy(rng.integers(0, n2, B))])
b = tr.to_device(b)
for _ in range(3): evaluate.full_catalogue_ranks(tr.P, tr.cfg, cat, b)
torch.cuda.synchronize()
_abi.profile(True)
t0 = time.perf_counter()
for _ in range(10): evaluate.full_catalogue_ranks(tr.P, tr.cfg, cat, b)
torch.cuda.synchronize()
print("ms per batch", (time.perf_counter() - t0) * 100)
rep = _abi.profile_report()
_abi.profile(False)
for k, (c, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1])[:8]: print(k, c, round(ms / 10, 4))

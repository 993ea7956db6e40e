import sys, os, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests/golden")
import torch, numpy as np
from amid_b200 import evaluate
from amid_b200.engine import Trainer
from amid_b200.model_seq import SASRec
V=894820; D=128; B=256; L=20; n1,n2=16084,12153
torch.manual_seed(2)
m = SASRec(0, D, V, D, L, 32, B, False, True, 0.5, 0.4).cuda().eval(); m.cfg.precision="x3"
tr = Trainer(m)
g = torch.Generator(device="cuda").manual_seed(9)
perm = torch.randperm(V, device="cuda", generator=g)[:n1+n2]
pool1, pool2 = perm[:n1].contiguous(), perm[n1:].contiguous()
cat = tr.catalogue(pool1, pool2)
ri = lambda hi,*s: torch.randint(0,hi,s,device="cuda",generator=g)
batches=[]
for _ in range(200):
    dom=ri(2,B)
    batches.append({"seq_d1":ri(V,B,L),"seq_d2":ri(V,B,L),"domain_id":dom,"i_node":torch.where(dom==0,pool1[ri(n1,B)],pool2[ri(n2,B)]),"overlap_label":ri(2,B)})
gf = evaluate.GraphedForward(tr.P, tr.cfg, B, L, 2, with_user_proj=True)
torch.cuda.synchronize()
for name, fn in (("replay only", lambda b: gf.graph.replay()), ("run (copies+replay)", lambda b: gf.run({"i_node":b["i_node"],"neg_samples":b["i_node"].view(B,1),"seq_d1":b["seq_d1"],"seq_d2":b["seq_d2"]}))):
    t0=time.perf_counter()
    for b in batches: fn(b)
    torch.cuda.synchronize()
    print(name, (time.perf_counter()-t0)/len(batches)*1e3, "ms/batch")
t0=time.perf_counter(); r=evaluate.evaluate_full_catalogue_fast(tr.P,tr.cfg,cat,batches,None); torch.cuda.synchronize(); print("fast total", (time.perf_counter()-t0)/len(batches)*1e3, "ms/batch")
import cProfile, pstats
pr=cProfile.Profile(); pr.enable(); evaluate.evaluate_full_catalogue_fast(tr.P,tr.cfg,cat,batches,None); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)

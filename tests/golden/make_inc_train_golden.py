"""Golden fixture F9: InnerComp + InterComp + DR heads in TRAINING direction (dropout p = 0), produced by executing the
reference model (model_seq.py:390-443 with isInC=True: positional table and attention over 2L, :399-402, 422-424) and
the phase-1 loss of train_sr_dr.py:217-221.  Pins the oracle's backward through inc_d*/itc_d* and the three heads.
Run once in the build container: python tests/golden/make_inc_train_golden.py"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (installs the shims, imports the reference modules)
from common import make_params  # noqa: E402


def main():
    d, hid, bs, L = 128, 32, 8, 6
    crit = nn.BCELoss(reduce=False)
    bt = G.draw_batches("cloth_sport_train75_DR.csv", bs, L, True, 199, 1, seed=9, dr=True)
    V, pad = G.compact(bt)
    P = make_params(19, V, d, 2 * L, hid, bs, isInC=True, isItC=True, isDR=True)
    ts1 = ts2 = 0.13
    m = G.build(P, V, d, L, hid, bs, True, True, ts1, ts2, True).train()
    for mod in m.modules():
        if isinstance(mod, nn.Dropout):
            mod.p = 0.0
        if isinstance(mod, nn.MultiheadAttention):
            mod.dropout = 0.0
    b = bt[0]
    lab, dom = b["label"].float(), b["domain_id"]
    m1, m2 = (1 - dom).unsqueeze(1), dom.unsqueeze(1)
    p1, p2, i1, i2, g1, g2 = G.fwd(m, b)
    lc = G.ref_loss_cls(crit, p1, p2, lab, dom)
    le = torch.mean((crit(p1, lab) - g1) ** 2 / i1 * m1 + (crit(p2, lab) - g2) ** 2 / i2 * m2)     # train_sr_dr.py:217-220
    loss = lc + le * 0.01                                                                              # :221
    m.zero_grad()
    loss.backward()
    arrs = {"V": V, "pad": pad, "ts": ts1, "loss_cls": lc.detach().numpy(), "loss_dr_e": le.detach().numpy(),
            **G.batch_np(b), **{k: v.detach().numpy() for k, v in dict(p1=p1, p2=p2, ips1=i1, ips2=i2, g1=g1, g2=g2).items()}}
    keep = ("inc_d", "itc_d", "predict")
    extra = ("sac1.pos_emb.weight", "sac2.last_layernorm.weight", "sac1.attention_layers.0.in_proj_weight",
             "sac2.forward_layers.1.conv2.weight", "sac1.attention_layernorms.1.bias")
    for n, prm in m.named_parameters():
        if n == "item_emb_layer.emb_item.weight":
            arrs["gtab_idx"], arrs["gtab_rows"] = G.sparse_rows(prm.grad)
        elif n.startswith(keep) or n in extra:
            arrs["grad/" + n] = prm.grad.numpy().copy()
    # the InnerComp gates of this batch must not all be closed, or the fixture pins nothing
    arrs["n_grad_tensors"] = len([k for k in arrs if k.startswith("grad/")])
    G.save("inc_train_small.npz", **arrs)
    print("inc_d1.trans_nn.weight grad max:", float(np.abs(arrs["grad/inc_d1.trans_nn.weight"]).max()))


if __name__ == "__main__":
    main()

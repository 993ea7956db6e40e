"""Golden fixture F10: the doubly-robust PHASE-2 step in training direction with dropout p = 0, produced by executing the
reference model (model_seq.py:390-443, isDR=True heads :403-416) and the phase-2 loss of train_sr_dr.py:392-394
(loss_dr_r, optimizer2).  dr_small.npz pins the same loss with injected dropout masks, which only the CPU oracle can
consume; this one lets the GPU path meet a reference-executed phase-2 loss and its gradients DIRECTLY.
Run once in the build container: python tests/golden/make_dr_phase2_golden.py"""
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
    d, hid, bs, L = 128, 32, 16, 20
    crit = nn.BCELoss(reduce=False)
    bt = G.draw_batches("cloth_sport_train75_DR.csv", bs, L, True, 199, 1, seed=11, dr=True)
    V, pad = G.compact(bt)
    P = make_params(23, V, d, L, hid, bs, isDR=True)
    ts2 = 0.07
    m = G.build(P, V, d, L, hid, bs, False, True, 0.5, ts2, True).train()
    for mod in m.modules():
        if isinstance(mod, nn.Dropout):
            mod.p = 0.0
        if isinstance(mod, nn.MultiheadAttention):
            mod.dropout = 0.0
    b = bt[0]
    lab, dom = b["label"].float(), b["domain_id"]
    m1, m2 = (1 - dom).unsqueeze(1), dom.unsqueeze(1)
    p1, p2, i1, i2, g1, g2 = G.fwd(m, b)
    ob = b["ob_label"].unsqueeze(1).repeat(1, 2)                                                       # train_sr_dr.py:392
    loss = torch.mean((g1 ** 2 + ob * ((crit(p1, lab) ** 2 - g1 ** 2) ** 2) / i1) * m1
                      + (g2 ** 2 + ob * ((crit(p2, lab) ** 2 - g2 ** 2) ** 2) / i2) * m2)               # :393-394
    m.zero_grad()
    loss.backward()
    arrs = {"V": V, "pad": pad, "ts": ts2, "loss_dr_r": loss.detach().numpy(), **G.batch_np(b),
            **{k: v.detach().numpy() for k, v in dict(p1=p1, p2=p2, ips1=i1, ips2=i2, g1=g1, g2=g2).items()}}
    n = 0
    for name, prm in m.named_parameters():
        if prm.grad is None:
            continue
        if name == "item_emb_layer.emb_item.weight":
            arrs["gtab_idx"], arrs["gtab_rows"] = G.sparse_rows(prm.grad)
        elif name.startswith(("predict", "itc_d")) or name in ("sac1.pos_emb.weight", "sac2.last_layernorm.weight",
                                                               "sac1.attention_layers.0.in_proj_weight",
                                                               "sac2.forward_layers.1.conv2.weight",
                                                               "sac1.attention_layernorms.1.bias"):
            arrs["grad/" + name] = prm.grad.numpy().copy()
            n += 1
    arrs["n_grad_tensors"] = n
    G.save("dr_phase2_nodrop.npz", **arrs)
    print("loss_dr_r", float(loss), "grad tensors", n, "gfunc grad max", float(np.abs(arrs["grad/predict_gfunc.fc.0.weight"]).max()))


if __name__ == "__main__":
    main()

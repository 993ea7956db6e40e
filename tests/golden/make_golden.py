"""Generate golden fixtures by EXECUTING THE REFERENCE (/root/reference) on CPU.

Run once in the build container:  python tests/golden/make_golden.py
Nothing at test / bench time reads /root/reference; only the .npz files written here.

Shims: (1) random.sample on a set -> tuple (Python >= 3.11);
(2) hard-coded .cuda()/device="cuda" -> CPU.  Dropout is made reproducible by replacing
torch.nn.functional.dropout with a version that consumes the keep-masks of
common.make_keep_masks in call order (the reference arithmetic `x * keep / (1-p)` is
unchanged).
"""
from __future__ import annotations

import os
import random
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
REF = "/root/reference"
sys.path.insert(0, REF)

from common import make_keep_masks, make_params  # noqa: E402

# ---------------------------------------------------------------- shims
_orig_sample = random.sample


def _sample(pop, k, **kw):
    if isinstance(pop, (set, frozenset)):
        pop = tuple(pop)
    return _orig_sample(pop, k, **kw)


random.sample = _sample
torch.Tensor.cuda = lambda self, *a, **k: self
nn.Module.cuda = lambda self, *a, **k: self
_orig_ones = torch.ones


def _ones(*a, **k):
    k.pop("device", None)
    return _orig_ones(*a, **k)


torch.ones = _ones

import dataset_seq  # noqa: E402
import model_seq  # noqa: E402
import utils as ref_utils  # noqa: E402

dataset_seq.random.sample = _sample

# ---------------------------------------------------------------- dropout injection
_orig_dropout = F.dropout
_mask_queue = []


def _dropout(x, p=0.5, training=True, inplace=False):
    if not training or p == 0.0:
        return x
    keep = _mask_queue.pop(0)
    assert keep.numel() == x.numel(), (keep.shape, x.shape)
    return x * keep.reshape(x.shape).to(x.dtype) * (1.0 / (1.0 - p))


def queue_masks(masks, B, L, d):
    """Order of F.dropout calls inside SASRec.forward (sac1 then sac2):
    emb [B,L,d]; per block: attn [B*H,L,L], ffn dropout1 [B,d,L], ffn dropout2 [B,d,L]."""
    _mask_queue.clear()
    for s in ("sac1", "sac2"):
        m = masks[s]
        _mask_queue.append(m["emb"])
        for i in range(2):
            _mask_queue.append(m[f"attn{i}"].reshape(B * 8, L, L))
            _mask_queue.append(m[f"ffn1_{i}"].transpose(1, 2).contiguous())
            _mask_queue.append(m[f"ffn2_{i}"].transpose(1, 2).contiguous())


F.dropout = _dropout
torch.nn.functional.dropout = _dropout

ITEM_LENGTH = 447410
PAD_ID = ITEM_LENGTH + 1


def draw_batches(csv, bs, L, is_train, neg_nums, n_batches, seed, dr=False, shuffle=True):
    torch.manual_seed(seed)
    np.random.seed(seed)
    random.seed(seed)
    cls = dataset_seq.DualDomainSeqDatasetDR if dr else dataset_seq.DualDomainSeqDataset
    coll = dataset_seq.collate_fn_enhanceDR if dr else dataset_seq.collate_fn_enhance
    ds = cls(seq_len=L, isTrain=is_train, neg_nums=neg_nums, long_length=7, pad_id=PAD_ID,
             csv_path=os.path.join(REF, "amazon_dataset", csv))
    dl = torch.utils.data.DataLoader(ds, batch_size=bs, shuffle=shuffle, num_workers=0, drop_last=True,
                                     collate_fn=coll)
    out = []
    for k, sample in enumerate(dl):
        if k >= n_batches:
            break
        out.append({key: v.long() for key, v in sample.items()})   # train_sr.py:191-199
    return out


def compact(batches):
    """Remap the raw item ids of a list of batches onto [0, V) so the table stays small."""
    ids = np.unique(np.concatenate([b[k].numpy().ravel() for b in batches
                                    for k in ("i_node", "neg_samples", "seq_d1", "seq_d2")]))
    lut = {int(v): i for i, v in enumerate(ids)}
    for b in batches:
        for k in ("i_node", "neg_samples", "seq_d1", "seq_d2"):
            b[k] = torch.from_numpy(np.vectorize(lut.get)(b[k].numpy()).astype(np.int64))
    return len(ids), lut.get(PAD_ID, -1)


def build(P, V, d, L, hid, bs, isInC, isItC, ts1, ts2, isDR):
    m = model_seq.SASRec(user_length=10, user_emb_dim=d, item_length=V, item_emb_dim=d, seq_len=L,
                         hid_dim=hid, bs=bs, isInC=isInC, isItC=isItC, threshold1=ts1, threshold2=ts2, isDR=isDR)
    missing = m.load_state_dict(P, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m


def fwd(m, b):
    return m(b["user_node"], b["i_node"], b["neg_samples"], b["seq_d1"], b["seq_d2"],
             b["long_tail_mask_d1"], b["long_tail_mask_d2"])


def batch_np(b):
    return {f"in_{k}": v.numpy() for k, v in b.items()}


def ref_loss_cls(crit, p1, p2, labels, domain_id):
    m1 = (1 - domain_id).unsqueeze(1)
    m2 = domain_id.unsqueeze(1)
    return torch.mean(crit(p1, labels) * m1 + crit(p2, labels) * m2)      # train_sr.py:210-211


def sparse_rows(t):
    nz = (t != 0).any(dim=1).nonzero().flatten()
    return nz.numpy(), t[nz].numpy()


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrs)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB")


def main():
    d, hid = 128, 32
    only = os.environ.get("GOLDEN_ONLY")
    crit = nn.BCELoss(reduce=False)

    # ---- F1: C1-shape eval-mode forward on a real cloth_sport_train75 batch (bs=256, L=20, ItC, ts2=0.4)
    bs, L = 256, 20
    bt = draw_batches("cloth_sport_train75.csv", bs, L, True, 199, 1, seed=0)
    V, pad = compact(bt)
    P = make_params(11, V, d, L, hid, bs)
    m = build(P, V, d, L, hid, bs, False, True, 0.5, 0.4, False).eval()
    feats = {}
    m.sac1.register_forward_hook(lambda mod, i, o: feats.__setitem__("enc1", o.detach().clone()))
    m.sac2.register_forward_hook(lambda mod, i, o: feats.__setitem__("enc2", o.detach().clone()))
    m.itc_d1.register_forward_hook(lambda mod, i, o: feats.__setitem__("f1", o.detach().clone()))
    m.itc_d2.register_forward_hook(lambda mod, i, o: feats.__setitem__("f2", o.detach().clone()))
    with torch.no_grad():
        p1, p2 = fwd(m, bt[0])
    lab = bt[0]["label"].float()
    save("c1_fwd_eval.npz", V=V, pad=pad, p1=p1.numpy(), p2=p2.numpy(),
         loss_cls=ref_loss_cls(crit, p1, p2, lab, bt[0]["domain_id"]).numpy(),
         enc1_head=feats["enc1"][:8].numpy(), enc2_head=feats["enc2"][:8].numpy(),
         E1=feats["f1"][0, L:].numpy(), E2=feats["f2"][0, L:].numpy(),
         u1=feats["f1"].mean(1).numpy(), u2=feats["f2"].mean(1).numpy(), **batch_np(bt[0]))

    # ---- F2: eval ranking on real cloth_sport_test batches, C = 100, incl. reference metrics
    C = 100
    bt = draw_batches("cloth_sport_test.csv", bs, L, False, C - 1, 3, seed=1, shuffle=False)
    V, pad = compact(bt)
    P = make_params(12, V, d, L, hid, bs)
    m = build(P, V, d, L, hid, bs, False, True, 0.5, 0.4, False).eval()
    p1s, p2s, doms, ovs = [], [], [], []
    with torch.no_grad():
        for b in bt:
            p1, p2 = fwd(m, b)
            p1s.append(p1.numpy().copy()); p2s.append(p2.numpy().copy())
            doms.append(b["domain_id"].numpy()); ovs.append(b["overlap_label"].numpy())
    p1a, p2a = np.concatenate(p1s), np.concatenate(p2s)
    dom, ov = np.concatenate(doms), np.concatenate(ovs)
    domx = np.repeat(dom[:, None], C, 1); ovx = np.repeat(ov[:, None], C, 1)
    l1, l2 = ref_utils.choose_predict(p1a, p2a, domx)
    o1, n1, o2, n2 = ref_utils.choose_predict_overlap(p1a, p2a, domx, ovx)
    met = {}
    for k, lst in (("d1_ov", o1), ("d1_no", n1), ("d2_ov", o2), ("d2_no", n2)):
        met["met_" + k] = np.array(ref_utils.get_sample_scores(lst), dtype=np.float64)
    l1 = l1.copy(); l2 = l2.copy()
    l1[:, 0] = l1[:, 0] - 1e-7; l2[:, 0] = l2[:, 0] - 1e-7               # train_sr.py:114-115
    met["met_d1"] = np.array(ref_utils.get_sample_scores(l1), dtype=np.float64)
    met["met_d2"] = np.array(ref_utils.get_sample_scores(l2), dtype=np.float64)
    met["ranks_d1"] = (-l1).argsort().argsort()[:, 0]
    met["ranks_d2"] = (-l2).argsort().argsort()[:, 0]
    arrs = {f"b{i}_{k}": v.numpy() for i, b in enumerate(bt) for k, v in b.items()
            if k in ("i_node", "neg_samples", "seq_d1", "seq_d2", "domain_id", "overlap_label")}
    save("c1_eval_rank.npz", V=V, pad=pad, p1=p1a, p2=p2a, **met, **arrs)

    # ---- F2b: ranking with ties / saturated scores (pure utils.py path)
    rng = np.random.default_rng(5)
    sc = rng.random((64, 50)).astype(np.float32)
    sc[:, 5:9] = sc[:, :1]                 # exact ties with the positive
    sc[10:20] = 1.0                        # fully saturated rows
    sc[20:30, 1:] = np.round(sc[20:30, 1:], 1)
    save("rank_ties.npz", scores=sc, ranks=(-sc).argsort().argsort()[:, 0],
         met=np.array(ref_utils.get_sample_scores(sc), dtype=np.float64))

    # ---- F3: train-mode (dropout masks injected) fwd + loss + all grads, then 3 Adam steps (bs=16)
    bs, L = 16, 20
    bt = draw_batches("cloth_sport_train75.csv", bs, L, True, 199, 3, seed=2)
    V, pad = compact(bt)
    P = make_params(13, V, d, L, hid, bs)
    m = build(P, V, d, L, hid, bs, False, True, 0.5, 0.07, False).train()
    opt = torch.optim.Adam(m.parameters(), lr=5e-4)
    arrs = {"V": V, "pad": pad}
    for step, b in enumerate(bt):
        queue_masks(make_keep_masks(100 + step, bs, L, d), bs, L, d)
        p1, p2 = fwd(m, b)
        assert not _mask_queue
        loss = ref_loss_cls(crit, p1, p2, b["label"].float(), b["domain_id"])
        opt.zero_grad(); loss.backward()
        if step == 0:
            arrs.update(p1=p1.detach().numpy(), p2=p2.detach().numpy(), loss=loss.detach().numpy())
            for n, prm in m.named_parameters():
                if n == "item_emb_layer.emb_item.weight":
                    arrs["gtab_idx"], arrs["gtab_rows"] = sparse_rows(prm.grad)
                else:
                    arrs["grad/" + n] = prm.grad.numpy().copy()
        opt.step()
        arrs[f"loss_step{step}"] = loss.detach().numpy()
        arrs.update({f"b{step}_{k}": v.numpy() for k, v in b.items()})
    for n, prm in m.named_parameters():
        arrs["after3/" + n] = prm.detach().numpy().copy()
    save("train_small.npz", **arrs)

    # ---- F4: DR (isDR) phase-1 and phase-2 losses + selected grads (bs=16), real _DR batch
    bt = draw_batches("cloth_sport_train75_DR.csv", bs, L, True, 199, 1, seed=3, dr=True)
    V, pad = compact(bt)
    P = make_params(14, V, d, L, hid, bs, isDR=True)
    m = build(P, V, d, L, hid, bs, False, True, 0.5, 0.07, True).train()
    b = bt[0]
    lab = b["label"].float()
    dom = b["domain_id"]
    m1 = (1 - dom).unsqueeze(1); m2 = dom.unsqueeze(1)
    sel = ["predictModule.fc.0.weight", "predict_ips.fc.0.weight", "predict_ips.fc.2.bias",
           "predict_gfunc.fc.0.weight", "predict_gfunc.fc.2.weight", "itc_d1.trans_bs.weight",
           "itc_d2.trans_nn.weight", "sac1.pos_emb.weight", "sac2.attention_layers.0.in_proj_weight",
           "sac1.forward_layers.1.conv1.weight", "sac2.last_layernorm.weight"]
    arrs = {"V": V, "pad": pad, **batch_np(b)}
    for phase in (1, 2):
        queue_masks(make_keep_masks(200, bs, L, d), bs, L, d)
        p1, p2, i1, i2, g1, g2 = fwd(m, b)
        if phase == 1:                                                        # train_sr_dr.py:217-221
            lc = ref_loss_cls(crit, p1, p2, lab, dom)
            le = torch.mean((crit(p1, lab) - g1) ** 2 / i1 * m1 + (crit(p2, lab) - g2) ** 2 / i2 * m2)
            loss = lc + le * 0.01
            arrs.update(loss_cls=lc.detach().numpy(), loss_dr_e=le.detach().numpy(),
                        **{k: v.detach().numpy() for k, v in
                           dict(p1=p1, p2=p2, ips1=i1, ips2=i2, g1=g1, g2=g2).items()})
        else:                                                                 # train_sr_dr.py:392-394
            ob = b["ob_label"].unsqueeze(1).repeat(1, 2)
            loss = torch.mean((g1 ** 2 + ob * ((crit(p1, lab) ** 2 - g1 ** 2) ** 2) / i1) * m1
                              + (g2 ** 2 + ob * ((crit(p2, lab) ** 2 - g2 ** 2) ** 2) / i2) * m2)
            arrs["loss_dr_r"] = loss.detach().numpy()
        m.zero_grad(); loss.backward()
        for n, prm in m.named_parameters():
            if n in sel:
                arrs[f"grad{phase}/" + n] = prm.grad.numpy().copy()
            if n == "item_emb_layer.emb_item.weight":
                arrs[f"gtab{phase}_idx"], arrs[f"gtab{phase}_rows"] = sparse_rows(prm.grad)
    save("dr_small.npz", **arrs)

    # ---- F5: InterComp / InnerComp literal outputs on inputs where some gates fire
    bsm, n = 12, 6
    g = torch.Generator().manual_seed(7)
    a = torch.randn(bsm, n, d, generator=g) * 0.2
    bb = torch.randn(bsm, n, d, generator=g) * 0.2
    a[3] *= 3.0; bb[3] *= 3.0; a[7] *= 3.0; bb[7] = a[7].clone()          # peaked users
    itc = model_seq.InterComp(d, bsm, 0.2)
    inc = model_seq.InnerComp(d, bsm, 0.2)
    PP = make_params(15, 4, d, n, hid, bsm, isInC=True)
    itc.load_state_dict({k[len("itc_d1."):]: v for k, v in PP.items() if k.startswith("itc_d1.")})
    inc.load_state_dict({k[len("inc_d1."):]: v for k, v in PP.items() if k.startswith("inc_d1.")})
    a.requires_grad_(True); bb.requires_grad_(True)
    o_itc = itc(a, bb)
    wgt = torch.randn(o_itc.shape, generator=g)
    (o_itc * wgt).sum().backward()
    arrs = dict(a=a.detach().numpy(), b=bb.detach().numpy(), itc_out=o_itc.detach().numpy(), wgt=wgt.numpy(),
                ga=a.grad.numpy().copy(), gb=bb.grad.numpy().copy(),
                **{"gitc/" + k: v.grad.numpy().copy() for k, v in itc.named_parameters()})
    a.grad = None
    o_inc = inc(a)
    (o_inc * wgt).sum().backward()
    arrs.update(inc_out=o_inc.detach().numpy(), ga_inc=a.grad.numpy().copy(),
                **{"ginc/" + k: v.grad.numpy().copy() for k, v in inc.named_parameters()})
    save("mim_peaked.npz", **arrs)

    # ---- F6: isInC + isItC eval forward (encoder length 2L), bs=16, L=10
    bs, L = 16, 10
    bt = draw_batches("cloth_sport_train75.csv", bs, L, True, 199, 1, seed=4)
    V, pad = compact(bt)
    P = make_params(16, V, d, 2 * L, hid, bs, isInC=True)
    m = build(P, V, d, L, hid, bs, True, True, 0.07, 0.07, False).eval()
    with torch.no_grad():
        p1, p2 = fwd(m, bt[0])
    save("inc_small.npz", V=V, pad=pad, p1=p1.numpy(), p2=p2.numpy(), **batch_np(bt[0]))

    # ---- F7: timeline mask actually firing (zero pad row + zero position rows), eval, bs=16, L=20
    bs, L = 16, 20
    bt = draw_batches("cloth_sport_train75.csv", bs, L, True, 199, 1, seed=5)
    V, pad = compact(bt)
    P = make_params(17, V, d, L, hid, bs, zero_rows=(pad,), zero_pos=(0, 1, 2, 5))
    m = build(P, V, d, L, hid, bs, False, True, 0.5, 0.07, False).eval()
    feats = {}
    m.sac1.register_forward_hook(lambda mod, i, o: feats.__setitem__("enc1", o.detach().clone()))
    with torch.no_grad():
        p1, p2 = fwd(m, bt[0])
    save("tmask.npz", V=V, pad=pad, p1=p1.numpy(), p2=p2.numpy(), enc1=feats["enc1"].numpy(), **batch_np(bt[0]))

    # ---- F8: train mode with dropout DISABLED (p = 0): fwd, loss, grads, 3 Adam steps -- directly
    #      comparable with the CUDA path (whose dropout RNG differs from torch's), bs=16, L=20
    bs, L = 16, 20
    bt = draw_batches("cloth_sport_train75.csv", bs, L, True, 199, 3, seed=6)
    V, pad = compact(bt)
    P = make_params(18, V, d, L, hid, bs)
    m = build(P, V, d, L, hid, bs, False, True, 0.5, 0.07, False).train()
    for mod in m.modules():
        if isinstance(mod, nn.Dropout):
            mod.p = 0.0
        if isinstance(mod, nn.MultiheadAttention):
            mod.dropout = 0.0
    opt = torch.optim.Adam(m.parameters(), lr=5e-4)
    sel8 = ["predictModule.fc.0.weight", "predictModule.fc.0.bias", "predictModule.fc.2.weight",
            "predictModule.fc.2.bias", "itc_d1.trans_bs.weight", "itc_d1.trans_bs.bias", "itc_d2.trans_nn.weight",
            "itc_d2.trans_nn.bias", "sac1.pos_emb.weight", "sac1.attention_layers.0.in_proj_weight",
            "sac1.attention_layers.0.in_proj_bias", "sac2.attention_layers.1.out_proj.weight",
            "sac2.attention_layers.1.out_proj.bias", "sac1.forward_layers.1.conv1.weight",
            "sac2.forward_layers.0.conv2.weight", "sac2.forward_layers.0.conv2.bias",
            "sac1.attention_layernorms.0.weight", "sac1.attention_layernorms.1.bias",
            "sac2.forward_layernorms.0.weight", "sac2.forward_layernorms.1.bias", "sac2.last_layernorm.weight",
            "sac1.last_layernorm.bias"]
    arrs = {"V": V, "pad": pad}
    for step, b in enumerate(bt):
        p1, p2 = fwd(m, b)
        loss = ref_loss_cls(crit, p1, p2, b["label"].float(), b["domain_id"])
        opt.zero_grad(); loss.backward()
        if step == 0:
            arrs.update(p1=p1.detach().numpy(), p2=p2.detach().numpy())
            for n, prm in m.named_parameters():
                if n == "item_emb_layer.emb_item.weight":
                    arrs["gtab_idx"], arrs["gtab_rows"] = sparse_rows(prm.grad)
                elif n in sel8:
                    arrs["grad/" + n] = prm.grad.numpy().copy()
        opt.step()
        arrs[f"loss_step{step}"] = loss.detach().numpy()
        arrs.update({f"b{step}_{k}": v.numpy() for k, v in b.items()})
    for n, prm in m.named_parameters():
        if n == "item_emb_layer.emb_item.weight":
            arrs["after3/" + n] = prm.detach().numpy().copy()
        elif n in sel8:
            arrs["after3/" + n] = prm.detach().numpy().copy()
    save("train_p0.npz", **arrs)


if __name__ == "__main__":
    main()

"""Pin the CPU oracle (oracle/amid_oracle.py) against fixtures produced by running the
reference itself (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import torch

from common import load, make_keep_masks, make_params
from oracle import amid_oracle as O

D, HID = 128, 32
T = lambda a: torch.from_numpy(np.asarray(a))


def _fwd(P, z, pre="in_", **kw):
    return O.sasrec_forward(P, T(z[pre + "i_node"]), T(z[pre + "neg_samples"]), T(z[pre + "seq_d1"]),
                            T(z[pre + "seq_d2"]), **kw)


def test_c1_forward_eval_closed_and_literal():
    z = load("c1_fwd_eval.npz")
    P = make_params(11, int(z["V"]), D, 20, HID, 256)
    col = {}
    p1, p2 = _fwd(P, z, isInC=False, isItC=True, ts1=0.5, ts2=0.4, isDR=False, collect=col)
    np.testing.assert_allclose(p1.numpy(), z["p1"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(p2.numpy(), z["p2"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(col["enc1"][:8].numpy(), z["enc1_head"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(col["enc2"][:8].numpy(), z["enc2_head"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(col["u1"].numpy(), z["u1"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(col["u2"].numpy(), z["u2"], rtol=0, atol=2e-5)
    lc = O.loss_cls(p1, p2, T(z["in_label"]).float(), T(z["in_domain_id"]))
    np.testing.assert_allclose(lc.numpy(), z["loss_cls"], rtol=1e-6)
    # gather is a bit-exact copy
    tab = P["item_emb_layer.emb_item.weight"]
    assert torch.equal(O.emb_gather(tab, T(z["in_seq_d1"])), tab[T(z["in_seq_d1"])])


def test_mim_closed_equals_reference_literal_with_gates_on():
    z = load("mim_peaked.npz")
    PP = make_params(15, 4, D, 6, HID, 12, isInC=True)
    a = T(z["a"]).requires_grad_(True)
    b = T(z["b"]).requires_grad_(True)
    w = {k: PP["itc_d1." + k].clone().requires_grad_(True) for k in
         ("trans_nn.weight", "trans_nn.bias", "trans_bs.weight", "trans_bs.bias")}
    col = {}
    out = O.mim_closed(a, b, w["trans_nn.weight"], w["trans_nn.bias"], w["trans_bs.weight"], w["trans_bs.bias"],
                       0.2, collect=col)
    assert 0 < col["g"].sum() < 12, "fixture must have some gates on and some off"
    np.testing.assert_allclose(out.detach().numpy(), z["itc_out"], rtol=0, atol=2e-6)
    lit = O.mim_literal(a, b, w["trans_nn.weight"], w["trans_nn.bias"], w["trans_bs.weight"], w["trans_bs.bias"], 0.2)
    np.testing.assert_allclose(lit.detach().numpy(), z["itc_out"], rtol=0, atol=2e-6)
    (out * T(z["wgt"])).sum().backward()
    np.testing.assert_allclose(a.grad.numpy(), z["ga"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(b.grad.numpy(), z["gb"], rtol=1e-5, atol=2e-5)
    for k, v in w.items():
        np.testing.assert_allclose(v.grad.numpy(), z["gitc/" + k], rtol=1e-4, atol=5e-5)
    # InnerComp = same op with both arguments the same sequence
    a2 = T(z["a"]).requires_grad_(True)
    out2 = O.mim_closed(a2, a2, PP["inc_d1.trans_nn.weight"], PP["inc_d1.trans_nn.bias"],
                        PP["inc_d1.trans_bs.weight"], PP["inc_d1.trans_bs.bias"], 0.2)
    np.testing.assert_allclose(out2.detach().numpy(), z["inc_out"], rtol=0, atol=2e-6)
    (out2 * T(z["wgt"])).sum().backward()
    np.testing.assert_allclose(a2.grad.numpy(), z["ga_inc"], rtol=1e-5, atol=2e-5)


def test_train_mode_grads_and_adam_trajectory():
    z = load("train_small.npz")
    V = int(z["V"])
    P = {k: v.clone().requires_grad_(True) for k, v in make_params(13, V, D, 20, HID, 16).items()}
    st = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in P.items()}
    for step in range(3):
        masks = make_keep_masks(100 + step, 16, 20, D)
        col = {}
        p1, p2 = _fwd(P, z, pre=f"b{step}_", isInC=False, isItC=True, ts1=0.5, ts2=0.07, isDR=False,
                      masks=masks, collect=col)
        loss = O.loss_cls(p1, p2, T(z[f"b{step}_label"]).float(), T(z[f"b{step}_domain_id"]))
        np.testing.assert_allclose(loss.detach().numpy(), z[f"loss_step{step}"], rtol=2e-5)
        for v in P.values():
            v.grad = None
        loss.backward()
        if step == 0:
            np.testing.assert_allclose(p1.detach().numpy(), z["p1"], rtol=0, atol=2e-6)
            for k, v in P.items():
                if k == "item_emb_layer.emb_item.weight":
                    gt = torch.zeros_like(v)
                    gt[T(z["gtab_idx"])] = T(z["gtab_rows"])
                    np.testing.assert_allclose(v.grad.numpy(), gt.numpy(), rtol=1e-4, atol=2e-7)
                else:
                    g = z["grad/" + k]
                    np.testing.assert_allclose(v.grad.numpy(), g, rtol=1e-3, atol=1e-6 + 1e-4 * np.abs(g).max(),
                                               err_msg=k)
        with torch.no_grad():
            for k, v in P.items():
                O.adam_step(v, v.grad, st[k][0], st[k][1], step + 1, 5e-4)
    for k, v in P.items():
        np.testing.assert_allclose(v.detach().numpy(), z["after3/" + k], rtol=0, atol=3e-5, err_msg=k)


def test_dr_losses_and_grads():
    z = load("dr_small.npz")
    V = int(z["V"])
    lab, dom, ob = T(z["in_label"]).float(), T(z["in_domain_id"]), T(z["in_ob_label"])
    for phase in (1, 2):
        P = {k: v.clone().requires_grad_(True) for k, v in make_params(14, V, D, 20, HID, 16, isDR=True).items()}
        p1, p2, i1, i2, g1, g2 = _fwd(P, z, isInC=False, isItC=True, ts1=0.5, ts2=0.07, isDR=True,
                                      masks=make_keep_masks(200, 16, 20, D))
        if phase == 1:
            for n, t in dict(p1=p1, p2=p2, ips1=i1, ips2=i2, g1=g1, g2=g2).items():
                np.testing.assert_allclose(t.detach().numpy(), z[n], rtol=0, atol=2e-6, err_msg=n)
            lc = O.loss_cls(p1, p2, lab, dom)
            le = O.loss_dr_e(p1, p2, i1, i2, g1, g2, lab, dom)
            np.testing.assert_allclose(lc.detach().numpy(), z["loss_cls"], rtol=2e-5)
            np.testing.assert_allclose(le.detach().numpy(), z["loss_dr_e"], rtol=2e-5)
            loss = lc + 0.01 * le
        else:
            loss = O.loss_dr_r(p1, p2, i1, i2, g1, g2, lab, dom, ob)
            np.testing.assert_allclose(loss.detach().numpy(), z["loss_dr_r"], rtol=2e-5)
        loss.backward()
        for k in z:
            if k.startswith(f"grad{phase}/"):
                g = z[k]
                np.testing.assert_allclose(P[k.split("/", 1)[1]].grad.numpy(), g, rtol=1e-3,
                                           atol=1e-7 + 1e-4 * np.abs(g).max(), err_msg=k)
        gt = torch.zeros(V, D)
        gt[T(z[f"gtab{phase}_idx"])] = T(z[f"gtab{phase}_rows"])
        np.testing.assert_allclose(P["item_emb_layer.emb_item.weight"].grad.numpy(), gt.numpy(), rtol=1e-3,
                                   atol=1e-7 + 1e-4 * gt.abs().max().item())


def test_inc_and_timeline_mask_variants():
    z = load("inc_small.npz")
    P = make_params(16, int(z["V"]), D, 20, HID, 16, isInC=True)
    p1, p2 = _fwd(P, z, isInC=True, isItC=True, ts1=0.07, ts2=0.07, isDR=False)
    np.testing.assert_allclose(p1.numpy(), z["p1"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(p2.numpy(), z["p2"], rtol=0, atol=2e-6)
    z = load("tmask.npz")
    P = make_params(17, int(z["V"]), D, 20, HID, 16, zero_rows=(int(z["pad"]),), zero_pos=(0, 1, 2, 5))
    col = {}
    p1, p2 = _fwd(P, z, isInC=False, isItC=True, ts1=0.5, ts2=0.07, isDR=False, collect=col)
    x0 = col["sac1"]["x0"]
    assert (x0 == 0).all(-1).any(), "fixture must contain fully-masked positions"
    np.testing.assert_allclose(col["enc1"].numpy(), z["enc1"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(p1.numpy(), z["p1"], rtol=0, atol=2e-6)


def test_rank_metrics_bit_exact():
    z = load("rank_ties.npz")
    r = O.rank_of_positive(z["scores"])
    assert np.array_equal(r, z["ranks"])
    assert O.metrics_from_ranks(r) == tuple(z["met"].tolist())
    z = load("c1_eval_rank.npz")
    dom = np.concatenate([z[f"b{i}_domain_id"] for i in range(3)])
    ov = np.concatenate([z[f"b{i}_overlap_label"] for i in range(3)])
    res = O.eval_lists(z["p1"], z["p2"], dom, ov)
    for k in ("d1", "d2", "d1_ov", "d1_no", "d2_ov", "d2_no"):
        assert res[k] == tuple(z["met_" + k].tolist()), k


def test_c1_eval_scores_on_test_batches():
    z = load("c1_eval_rank.npz")
    P = make_params(12, int(z["V"]), D, 20, HID, 256)
    for i in range(3):
        p1, p2 = _fwd(P, z, pre=f"b{i}_", isInC=False, isItC=True, ts1=0.5, ts2=0.4, isDR=False)
        np.testing.assert_allclose(p1.numpy(), z["p1"][i * 256:(i + 1) * 256], rtol=0, atol=3e-6)
        np.testing.assert_allclose(p2.numpy(), z["p2"][i * 256:(i + 1) * 256], rtol=0, atol=3e-6)


def test_train_p0_trajectory():
    """Train mode with dropout disabled -- the fixture the CUDA path is compared with directly."""
    z = load("train_p0.npz")
    V = int(z["V"])
    P = {k: v.clone().requires_grad_(True) for k, v in make_params(18, V, D, 20, HID, 16).items()}
    st = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in P.items()}
    for step in range(3):
        p1, p2 = _fwd(P, z, pre=f"b{step}_", isInC=False, isItC=True, ts1=0.5, ts2=0.07, isDR=False)
        loss = O.loss_cls(p1, p2, T(z[f"b{step}_label"]).float(), T(z[f"b{step}_domain_id"]))
        np.testing.assert_allclose(loss.detach().numpy(), z[f"loss_step{step}"], rtol=2e-5)
        for v in P.values():
            v.grad = None
        loss.backward()
        if step == 0:
            for k in z:
                if k.startswith("grad/"):
                    g = z[k]
                    np.testing.assert_allclose(P[k[5:]].grad.numpy(), g, rtol=1e-3, atol=1e-7 + 1e-4 * np.abs(g).max(),
                                               err_msg=k)
        with torch.no_grad():
            for k, v in P.items():
                O.adam_step(v, v.grad, st[k][0], st[k][1], step + 1, 5e-4)
    for k in z:
        if k.startswith("after3/"):
            np.testing.assert_allclose(P[k[7:]].detach().numpy(), z[k], rtol=0, atol=3e-5, err_msg=k)


def test_inc_itc_dr_training_direction_golden():
    """F9 (tests/golden/make_inc_train_golden.py): InnerComp + InterComp + DR heads, dropout off, executed by the
    reference -- six outputs, both phase-1 losses and the gradients through inc_d*/itc_d*/predict* and the table."""
    z = load("inc_train_small.npz")
    V, ts = int(z["V"]), float(z["ts"])
    P = {k: v.clone().requires_grad_(True)
         for k, v in make_params(19, V, D, 12, HID, 8, isInC=True, isItC=True, isDR=True).items()}
    outs = _fwd(P, z, isInC=True, isItC=True, ts1=ts, ts2=ts, isDR=True)
    for n, t in zip(("p1", "p2", "ips1", "ips2", "g1", "g2"), outs):
        np.testing.assert_allclose(t.detach().numpy(), z[n], rtol=0, atol=2e-6, err_msg=n)
    lab, dom = T(z["in_label"]).float(), T(z["in_domain_id"])
    lc = O.loss_cls(outs[0], outs[1], lab, dom)
    le = O.loss_dr_e(*outs, lab, dom)
    np.testing.assert_allclose(lc.detach().numpy(), z["loss_cls"], rtol=2e-5)
    np.testing.assert_allclose(le.detach().numpy(), z["loss_dr_e"], rtol=2e-5)
    (lc + 0.01 * le).backward()
    n_checked = 0
    for k in z:
        if k.startswith("grad/"):
            g = z[k]
            np.testing.assert_allclose(P[k.split("/", 1)[1]].grad.numpy(), g, rtol=1e-3,
                                       atol=1e-7 + 1e-4 * np.abs(g).max(), err_msg=k)
            n_checked += 1
    assert n_checked == int(z["n_grad_tensors"]) and n_checked >= 20
    assert np.abs(z["grad/inc_d1.trans_nn.weight"]).max() > 0          # the InnerComp gates were open
    gt = torch.zeros(V, D)
    gt[T(z["gtab_idx"])] = T(z["gtab_rows"])
    got = P["item_emb_layer.emb_item.weight"].grad
    np.testing.assert_allclose(got.numpy(), gt.numpy(), rtol=1e-3, atol=1e-7 + 1e-4 * float(gt.abs().max()))


def test_dr_phase2_training_direction_golden():
    """F10 (tests/golden/make_dr_phase2_golden.py): the doubly-robust phase-2 loss (train_sr_dr.py:392-394) and its gradients,
    dropout off, executed by the reference -- the fixture the GPU path meets directly in tests/test_gpu_parity.py."""
    z = load("dr_phase2_nodrop.npz")
    V, ts = int(z["V"]), float(z["ts"])
    P = {k: v.clone().requires_grad_(True) for k, v in make_params(23, V, D, 20, HID, 16, isDR=True).items()}
    outs = _fwd(P, z, isInC=False, isItC=True, ts1=0.5, ts2=ts, isDR=True)
    for n, t in zip(("p1", "p2", "ips1", "ips2", "g1", "g2"), outs):
        np.testing.assert_allclose(t.detach().numpy(), z[n], rtol=0, atol=2e-6, err_msg=n)
    lab, dom, ob = T(z["in_label"]).float(), T(z["in_domain_id"]), T(z["in_ob_label"])
    loss = O.loss_dr_r(*outs, lab, dom, ob)
    np.testing.assert_allclose(loss.detach().numpy(), z["loss_dr_r"], rtol=2e-5)
    loss.backward()
    n_checked = 0
    for k in z:
        if k.startswith("grad/"):
            g = z[k]
            got = P[k.split("/", 1)[1]].grad
            got = torch.zeros_like(P[k.split("/", 1)[1]]) if got is None else got
            np.testing.assert_allclose(got.numpy(), g, rtol=1e-3, atol=1e-7 + 1e-4 * np.abs(g).max(), err_msg=k)
            n_checked += 1
    assert n_checked == int(z["n_grad_tensors"]) and n_checked >= 20
    gt = torch.zeros(V, D)
    gt[T(z["gtab_idx"])] = T(z["gtab_rows"])
    np.testing.assert_allclose(P["item_emb_layer.emb_item.weight"].grad.numpy(), gt.numpy(), rtol=1e-3,
                               atol=1e-7 + 1e-4 * float(gt.abs().max()))

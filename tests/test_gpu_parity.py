"""GPU parity tests: the CUDA hot path (through the C ABI) against the CPU oracle and the
golden fixtures produced by the reference.  Run on the B200 box with `-m gpu`.

Tolerances (fp32 path): probabilities / features agree with the reference to <= 2e-5 abs
(fp32 summation order differs between cuBLAS/MKL and the tile kernels); gradients to
2e-4 of the tensor's max magnitude; index/gather outputs and rankings are bit-exact.
"""
import numpy as np
import pytest
import torch

from helpers import (D, HID, O, T, assert_close, batch_from, build_model, grad_tol, load, make_params,
                     oracle_forward, random_batch, run_model, to_cuda)

pytestmark = pytest.mark.gpu

# the two parity-grade modes: exact fp32 CUDA-core tiles, and the split-operand tcgen05 path (same tolerances)
PRECS = ["fp32", "x3"]


def _hp():
    from amid_b200 import hotpath
    return hotpath


def _assert_trajectory(got, want, prec, lr, steps, msg):
    """Parameters after `steps` Adam steps.  fp32 tiles: every element within 3e-5.  Split-operand tiles carry 22-bit
    operands (the gradients agree to the same 2e-4 max|g| as fp32, checked separately); Adam divides by |g| + 1e-8, so
    an element whose gradient is within ~1e-9 of zero turns that last-bit noise into a step of up to lr.  Such
    elements may exceed 3e-5 -- at most 0.02 % of a tensor, never by more than steps * lr."""
    got = got.detach().cpu().numpy() if torch.is_tensor(got) else np.asarray(got)
    want = np.asarray(want)
    diff = np.abs(got - want)
    if prec == "fp32":
        assert diff.max() <= 3e-5, (msg, diff.max())
        return
    bad = diff > 3e-5
    assert bad.sum() <= max(1, int(2e-4 * diff.size)), (msg, int(bad.sum()), diff.size)
    assert diff.max() <= steps * lr * 1.01, (msg, diff.max())


# ------------------------------------------------------------------ a1 / a2: gather (bit-exact)
@pytest.mark.parametrize("n_rows", [1, 3, 4, 5, 33, 1000, 4099])
def test_gather_bit_exact(n_rows):
    hp = _hp()
    from amid_b200._abi import call, lib
    g = torch.Generator().manual_seed(n_rows)
    V = 5000
    table = torch.randn(V, D, generator=g).cuda()
    ids = torch.randint(0, V, (n_rows,), generator=g)
    ids[: n_rows // 2] = 7                     # duplicates / hot row
    ids = ids.cuda()
    out = torch.empty(n_rows, D, device="cuda")
    call("amid_emb_gather_fwd", hp._ptr(table), V, hp._ptr(ids), n_rows, hp._ptr(out), hp._stream())
    torch.cuda.synchronize()
    assert torch.equal(out, table[ids])
    assert lib().amid_gather_error_host_sync() == 0


def test_gather_rejects_bad_ids_and_empty():
    hp = _hp()
    from amid_b200._abi import call, lib
    table = torch.randn(10, D).cuda()
    ids = torch.tensor([1, 12, 3, -1], dtype=torch.long).cuda()
    out = torch.zeros(4, D, device="cuda")
    call("amid_emb_gather_fwd", hp._ptr(table), 10, hp._ptr(ids), 4, hp._ptr(out), hp._stream())
    assert lib().amid_gather_error_host_sync() == 1          # flagged, not silently wrapped
    assert torch.equal(out[0], table[1]) and torch.equal(out[2], table[3]) and out[1].abs().sum() == 0
    call("amid_emb_gather_fwd", hp._ptr(table), 10, hp._ptr(ids), 0, hp._ptr(out), hp._stream())   # empty: no-op


@pytest.mark.parametrize("B,L", [(1, 1), (3, 7), (16, 20), (5, 200)])
def test_seq_embed_bit_exact_and_mask_bits(B, L):
    import ctypes as C
    hp = _hp()
    from amid_b200._abi import Dropout, call
    g = torch.Generator().manual_seed(B * 1000 + L)
    V = 300
    table = torch.randn(V, D, generator=g)
    pos = torch.randn(L, D, generator=g)
    table[5] = 0.0
    pos[0] = 0.0                                 # (row 5, position 0) -> a fully masked position
    pos[L - 1, 3] = -table[9, 3]                 # a single masked element
    ids = torch.randint(0, V, (B, L), generator=g)
    ids[0, 0] = 5
    ids[B - 1, L - 1] = 9
    x0 = torch.empty(B * L, D, device="cuda")
    tm = torch.zeros(B * L * 4, dtype=torch.int32, device="cuda")
    drop = Dropout(0, 0.5, 0, 0)
    tc, pc, ic = table.cuda(), pos.cuda(), ids.cuda()
    call("amid_seq_embed_fwd", hp._ptr(tc), V, hp._ptr(ic), None, hp._ptr(pc), B, L, hp._ptr(x0), hp._ptr(tm),
         C.byref(drop), hp._stream())
    want = table[ids] + pos[:L]                  # model_seq.py:362, same fp32 add
    assert torch.equal(x0.view(B, L, D).cpu(), want)
    bits = tm.view(B * L, 4).cpu().numpy().astype(np.uint32)
    mask = np.zeros((B * L, D), dtype=bool)
    for e in range(4):
        for j in range(32):
            mask[:, 4 * j + e] = (bits[:, e] >> j) & 1
    assert np.array_equal(mask, (want == 0).view(B * L, D).numpy())
    if B * L > 1:
        assert mask[0].all() and mask[B * L - 1, 3]


@pytest.mark.parametrize("B,L,C", [(1, 1, 2), (3, 7, 5), (37, 20, 2), (9, 200, 3)])
def test_embed_all_single_launch_bit_exact(B, L, C):
    """The fused one-launch gather (candidates + both sequences with pos add and mask bits)."""
    import ctypes as Ct
    hp = _hp()
    from amid_b200._abi import Dropout, call
    g = torch.Generator().manual_seed(B * 131 + L)
    V = 1000
    table = torch.randn(V, D, generator=g)
    pos = [torch.randn(L, D, generator=g), torch.randn(L, D, generator=g)]
    ids_items = torch.randint(0, V, (B, C), generator=g)
    seqs = [torch.randint(0, V, (B, L), generator=g), torch.randint(0, V, (B, L), generator=g)]
    table[3] = 0.0
    pos[1][0] = 0.0
    seqs[1][0, 0] = 3
    tc = table.cuda()
    outs = [torch.empty(B * C, D, device="cuda"), torch.empty(B * L, D, device="cuda"), torch.empty(B * L, D, device="cuda")]
    tms = [torch.zeros(B * L * 4, dtype=torch.int32, device="cuda") for _ in range(2)]
    drop = Dropout(0, 0.5, 0, 0)
    ii, s1, s2, p1, p2 = ids_items.cuda(), seqs[0].cuda(), seqs[1].cuda(), pos[0].cuda(), pos[1].cuda()
    call("amid_embed_all_fwd", hp._ptr(tc), V, hp._ptr(ii), B * C, hp._ptr(s1), hp._ptr(s2), hp._ptr(p1), hp._ptr(p2), B, L,
         hp._ptr(outs[0]), hp._ptr(outs[1]), hp._ptr(outs[2]), hp._ptr(tms[0]), hp._ptr(tms[1]), Ct.byref(drop), hp._stream())
    assert torch.equal(outs[0].cpu(), table[ids_items.reshape(-1)])
    for k in range(2):
        want = (table[seqs[k]] + pos[k][:L]).view(B * L, D)
        assert torch.equal(outs[k + 1].cpu(), want)
        bits = tms[k].view(B * L, 4).cpu().numpy().astype(np.uint32)
        mask = np.zeros((B * L, D), dtype=bool)
        for e in range(4):
            for j in range(32):
                mask[:, 4 * j + e] = (bits[:, e] >> j) & 1
        assert np.array_equal(mask, (want == 0).numpy())
    assert (tms[1].view(B * L, 4)[0] != 0).all()


# ------------------------------------------------------------------ whole forward vs reference goldens
@pytest.mark.parametrize("prec", PRECS)
def test_forward_c1_golden(prec):
    z = load("c1_fwd_eval.npz")
    P = make_params(11, int(z["V"]), D, 20, HID, 256)
    m = build_model(P, int(z["V"]), 20, 256, ts2=0.4, precision=prec).eval()
    b = batch_from(z)
    with torch.no_grad():
        p1, p2 = run_model(m, b)
    assert_close(p1, z["p1"], 0, 2e-5)
    assert_close(p2, z["p2"], 0, 2e-5)
    hp = _hp()
    probs, ctx = hp.forward(m.param_dict(), m.cfg, b["i_node"], b["neg_samples"], b["seq_d1"], b["seq_d2"],
                            train=False)
    assert_close(ctx.encs[0].view(256, 20, D)[:8], z["enc1_head"], 0, 5e-5)
    assert_close(ctx.encs[1].view(256, 20, D)[:8], z["enc2_head"], 0, 5e-5)
    assert_close(ctx.us[0], z["u1"], 0, 5e-5)
    assert_close(ctx.us[1], z["u2"], 0, 5e-5)
    assert_close(ctx.itcs[0].E, z["E1"], 0, 5e-5)
    assert_close(ctx.itcs[1].E, z["E2"], 0, 5e-5)
    losses, _ = hp.loss_fwd_bwd(probs, b["label"], b["domain_id"], None, 0, 0.0, 256)
    assert_close(losses[0], z["loss_cls"], 2e-5, 0)


def test_forward_eval_batches_and_rank_golden():
    from amid_b200 import evaluate
    z = load("c1_eval_rank.npz")
    P = make_params(12, int(z["V"]), D, 20, HID, 256)
    m = build_model(P, int(z["V"]), 20, 256, ts2=0.4).eval()
    p1s, p2s = [], []
    with torch.no_grad():
        for i in range(3):
            p1, p2 = run_model(m, batch_from(z, pre=f"b{i}_"))
            p1s.append(p1), p2s.append(p2)
    p1, p2 = torch.cat(p1s), torch.cat(p2s)
    assert_close(p1, z["p1"], 0, 3e-5)
    assert_close(p2, z["p2"], 0, 3e-5)
    # rankings / metrics: bit-exact given identical scores (the reference's own scores)
    dom = torch.cat([T(z[f"b{i}_domain_id"]) for i in range(3)]).cuda()
    ov = torch.cat([T(z[f"b{i}_overlap_label"]) for i in range(3)]).cuda()
    res = evaluate.evaluate_lists(T(z["p1"]).cuda(), T(z["p2"]).cuda(), dom, ov)
    for k in ("d1", "d2", "d1_ov", "d1_no", "d2_ov", "d2_no"):
        assert res[k] == tuple(z["met_" + k].tolist()), k
    r1 = evaluate.rank_of_positive(T(z["p1"]).cuda()[dom == 0], evaluate.FIX_VALUE)
    assert np.array_equal(r1, z["ranks_d1"])
    # and end to end with our own scores: same rankings unless two scores are within fp32 noise
    ours = evaluate.evaluate_lists(p1, p2, dom, ov)
    for k in ("d1", "d2"):
        assert abs(ours[k][4] - z["met_" + k][4]) <= 2.0 / 256, k      # HIT@10 moves by at most a couple of users


def test_rank_with_ties_bit_exact():
    from amid_b200 import evaluate
    z = load("rank_ties.npz")
    r = evaluate.rank_of_positive(T(z["scores"]).cuda(), 0.0)
    assert np.array_equal(r, z["ranks"])
    assert evaluate.metrics_from_ranks(r) == tuple(z["met"].tolist())
    assert len(evaluate.rank_of_positive(torch.empty(0, 5, device="cuda"), 0.0)) == 0


@pytest.mark.parametrize("prec", PRECS)
def test_inc_and_tmask_goldens(prec):
    z = load("inc_small.npz")
    P = make_params(16, int(z["V"]), D, 20, HID, 16, isInC=True)
    m = build_model(P, int(z["V"]), 10, 16, isInC=True, ts1=0.07, ts2=0.07, precision=prec).eval()
    with torch.no_grad():
        p1, p2 = run_model(m, batch_from(z))
    assert_close(p1, z["p1"], 0, 2e-5)
    assert_close(p2, z["p2"], 0, 2e-5)
    z = load("tmask.npz")
    P = make_params(17, int(z["V"]), D, 20, HID, 16, zero_rows=(int(z["pad"]),), zero_pos=(0, 1, 2, 5))
    m = build_model(P, int(z["V"]), 20, 16, ts2=0.07, precision=prec).eval()
    b = batch_from(z)
    probs, ctx = _hp().forward(m.param_dict(), m.cfg, b["i_node"], b["neg_samples"], b["seq_d1"], b["seq_d2"],
                               train=False)
    assert_close(ctx.encs[0].view(16, 20, D), z["enc1"], 0, 5e-5)
    assert_close(probs[0, 0], z["p1"], 0, 2e-5)


# ------------------------------------------------------------------ train mode, dropout off: direct golden
@pytest.mark.parametrize("prec", PRECS)
def test_train_p0_grads_and_trajectory_golden(prec):
    from amid_b200.engine import Trainer
    z = load("train_p0.npz")
    V = int(z["V"])
    P = make_params(18, V, D, 20, HID, 16)
    # (1) drop-in autograd path, unchanged-driver style: BCELoss + backward through the module
    m = build_model(P, V, 20, 16, ts2=0.07, drop_p=0.0, precision=prec).train()
    b = batch_from(z, pre="b0_")
    p1, p2 = run_model(m, b)
    crit = torch.nn.BCELoss(reduction="none")
    dom = b["domain_id"]
    loss = torch.mean(crit(p1, b["label"]) * (1 - dom).unsqueeze(1) + crit(p2, b["label"]) * dom.unsqueeze(1))
    assert_close(loss, z["loss_step0"], 2e-5, 0)
    loss.backward()
    named = dict(m.named_parameters())
    for k in z:
        if k.startswith("grad/"):
            assert_close(named[k[5:]].grad, z[k], 1e-3, grad_tol(z[k]), k)
    gt = torch.zeros(V, D)
    gt[T(z["gtab_idx"])] = T(z["gtab_rows"])
    assert_close(named["item_emb_layer.emb_item.weight"].grad, gt, 1e-3, grad_tol(gt))
    # (2) fused engine: 3 Adam steps, sparse table update with exact dense semantics
    for sparse in (True, False):
        m2 = build_model(P, V, 20, 16, ts2=0.07, drop_p=0.0, precision=prec).train()
        tr = Trainer(m2, lr=5e-4, sparse_table=sparse)
        for step in range(3):
            losses = tr.step(batch_from(z, pre=f"b{step}_"))
            assert_close(losses[0], z[f"loss_step{step}"], 5e-5, 0, f"step {step}")
        tr.flush()
        named = dict(m2.named_parameters())
        for k in z:
            if k.startswith("after3/"):
                _assert_trajectory(named[k[7:]], z[k], prec, 5e-4, 3, f"{k} sparse={sparse}")


@pytest.mark.parametrize("prec", PRECS)
def test_dropin_unchanged_driver_loop_with_torch_adam(prec):
    """The reference training loop body (train_sr.py:191-215) driving the drop-in module with
    torch.optim.Adam over model.parameters(): 3 steps must land on the reference's parameters."""
    z = load("train_p0.npz")
    V = int(z["V"])
    model = build_model(make_params(18, V, D, 20, HID, 16), V, 20, 16, ts2=0.07, drop_p=0.0, precision=prec)
    optimizer = torch.optim.Adam(model.parameters(), lr=5e-4)            # train_sr.py:480
    criterion_cls = torch.nn.BCELoss(reduction="none")                   # train_sr.py:184
    model.train()
    for step in range(3):
        sample = {k[len(f"b{step}_"):]: T(z[k]) for k in z if k.startswith(f"b{step}_")}
        u_node = sample["user_node"].long().cuda()
        i_node = sample["i_node"].long().cuda()
        neg_samples = sample["neg_samples"].long().cuda()
        seq_d1 = sample["seq_d1"].long().cuda()
        seq_d2 = sample["seq_d2"].long().cuda()
        lt1 = sample["long_tail_mask_d1"].long().cuda()
        lt2 = sample["long_tail_mask_d2"].long().cuda()
        domain_id = sample["domain_id"].long().cuda()
        labels = sample["label"].long().cuda().float()
        predict_d1, predict_d2 = model(u_node, i_node, neg_samples, seq_d1, seq_d2, lt1, lt2)
        predict_d1, predict_d2 = predict_d1.squeeze(), predict_d2.squeeze()
        mask_d1 = (torch.ones_like(domain_id) - domain_id)
        mask_d2 = domain_id
        loss_cls = criterion_cls(predict_d1, labels) * mask_d1.unsqueeze(1) + criterion_cls(predict_d2, labels) * mask_d2.unsqueeze(1)
        loss = torch.mean(loss_cls)
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        assert_close(loss, z[f"loss_step{step}"], 5e-5, 0, f"step {step}")
    named = dict(model.named_parameters())
    for k in z:
        if k.startswith("after3/"):
            _assert_trajectory(named[k[7:]], z[k], prec, 5e-4, 3, k)
    sd = model.state_dict()
    assert sd["item_emb_layer.emb_item.weight"].shape == (V, D)


@pytest.mark.parametrize("prec", PRECS)
def test_trainer_dr_two_phase_two_adams_vs_oracle(prec):
    """train_sr_dr.py: phase 1 (loss_cls + dr_e_w * loss_dr_e, `optimizer`) and phase 2 (loss_dr_r,
    `optimizer2` with lr * lr2) interleave on the same parameters, each Adam with its own state and
    dense semantics on the table."""
    from amid_b200.engine import Trainer
    B, L, C, V = 8, 12, 2, 53
    rng = np.random.default_rng(21)
    P = make_params(33, V, D, L, HID, B, isDR=True)
    m = build_model(P, V, L, B, ts2=0.3, isDR=True, drop_p=0.0, precision=prec).train()
    tr = Trainer(m, lr=5e-4, lr2=0.5, dr_e_w=0.01)
    Po = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    st = [{k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in Po.items()} for _ in range(2)]
    steps = [0, 0]
    for phase in (1, 2, 2, 1, 2, 1):
        b = random_batch(rng, B, L, C, V)
        losses = tr.step(to_cuda(b), phase=phase)
        outs = oracle_forward(Po, b, isInC=False, isItC=True, ts1=0.5, ts2=0.3, isDR=True)
        lab, dom, ob = b["label"], b["domain_id"], b["ob_label"]
        if phase == 1:
            lc, le = O.loss_cls(outs[0], outs[1], lab, dom), O.loss_dr_e(*outs, lab, dom)
            lo = lc + 0.01 * le
            assert_close(losses[0], lc, 3e-5, 0)
            assert_close(losses[1], le, 3e-5, 1e-7)
        else:
            lo = O.loss_dr_r(*outs, lab, dom, ob)
            assert_close(losses[2], lo, 3e-5, 1e-7)
        for v in Po.values():
            v.grad = None
        lo.backward()
        oi = phase - 1
        steps[oi] += 1
        with torch.no_grad():
            for k, v in Po.items():
                g = v.grad if v.grad is not None else torch.zeros_like(v)
                O.adam_step(v, g, st[oi][k][0], st[oi][k][1], steps[oi], 5e-4 if oi == 0 else 2.5e-4)
    tr.flush()
    named = dict(m.named_parameters())
    for k, v in Po.items():
        assert_close(named[k], v, 0, 3e-5, k)


# ------------------------------------------------------------------ C3 sequence length, several CTA waves: oracle + our masks
@pytest.mark.parametrize("prec", PRECS)
def test_c3_scale_train_step_vs_oracle(prec):
    """B = 64, L = 200 (the C3 sequence length; 512 (sample, head) attention CTAs = several waves, 100 token tiles per chain
    kernel), dropout ON with the masks the kernels export, ItC gates firing: probabilities, loss and every gradient against
    the oracle at the fp32 tolerances, in both parity-grade modes."""
    hp = _hp()
    B, L, C, V, ts2 = 64, 200, 2, 997, 0.0155
    rng = np.random.default_rng(20241)
    P = make_params(57, V, D, L, HID, B)
    m = build_model(P, V, L, B, ts2=ts2, precision=prec).train()
    b = to_cuda(random_batch(rng, B, L, C, V))
    seed = 424242
    probs, ctx = hp.forward(m.param_dict(), m.cfg, b["i_node"], b["neg_samples"], b["seq_d1"], b["seq_d2"], train=True, seed=seed)
    masks = {s: {k: v.cpu() for k, v in d.items()} for s, d in hp.dropout_masks(m.cfg, B, L, seed, "cuda").items()}
    Po = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    col = {}
    outs = oracle_forward(Po, b, isInC=False, isItC=True, ts1=0.5, ts2=ts2, isDR=False, masks=masks, collect=col)
    pj = torch.softmax(O.mim_scores(col["enc1"], col["enc2"]), 0)
    assert 0 < int((pj > ts2).sum()) < B, "the case is meant to have some, not all, ItC gates open"
    if (pj - ts2).abs().min() < 1e-5:
        pytest.skip("gate margin too small for a meaningful comparison")
    for i, o in enumerate(outs):
        assert_close(probs[i // 2, i % 2], o, 0, 3e-5, f"out {i}")
    losses, dprobs = hp.loss_fwd_bwd(probs, b["label"], b["domain_id"], b["ob_label"], 0, 0.01, B)
    lo = O.loss_cls(outs[0], outs[1], b["label"].cpu(), b["domain_id"].cpu())
    assert_close(losses[0], lo, 3e-5, 0)
    lo.backward()
    G, ids_all, rows_all = hp.backward(m.param_dict(), m.cfg, ctx, dprobs)
    for k, v in Po.items():
        if k == "item_emb_layer.emb_item.weight":
            uid, ug, nu = hp.segreduce(ids_all, rows_all, V)
            assert_close(hp.dense_table_grad(uid, ug, nu, V), v.grad, 1e-3, grad_tol(v.grad), k)
        elif v.grad is not None:
            assert_close(G[k], v.grad, 1e-3, grad_tol(v.grad), k)


# ------------------------------------------------------------------ train mode WITH dropout: oracle + our masks
@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("B,L,C,isDR", [(6, 9, 2, False), (5, 20, 3, True)])
def test_train_dropout_vs_oracle_with_exported_masks(B, L, C, isDR, prec):
    hp = _hp()
    rng = np.random.default_rng(B * 100 + L)
    V = 97
    P = make_params(31, V, D, L, HID, B, isDR=isDR)
    m = build_model(P, V, L, B, ts2=0.2, isDR=isDR, precision=prec).train()
    b = to_cuda(random_batch(rng, B, L, C, V))
    seed = 1234567
    probs, ctx = hp.forward(m.param_dict(), m.cfg, b["i_node"], b["neg_samples"], b["seq_d1"], b["seq_d2"],
                            train=True, seed=seed)
    masks = hp.dropout_masks(m.cfg, B, L, seed, "cuda")
    masks = {s: {k: v.cpu() for k, v in d.items()} for s, d in masks.items()}
    keep_rate = float(np.mean([v.float().mean().item() for d in masks.values() for v in d.values()]))
    assert 0.45 < keep_rate < 0.55
    Po = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    col = {}
    outs = oracle_forward(Po, b, isInC=False, isItC=True, ts1=0.5, ts2=0.2, isDR=isDR, masks=masks, collect=col)
    pj = torch.softmax(O.mim_scores(col["enc1"], col["enc2"]), 0)
    if (pj - 0.2).abs().min() < 1e-4:
        pytest.skip("gate margin too small for a meaningful comparison")
    for i, o in enumerate(outs):
        assert_close(probs[i // 2, i % 2], o, 0, 3e-5, f"out {i}")
    lab, dom, ob = b["label"].cpu(), b["domain_id"].cpu(), b["ob_label"].cpu()
    modes = [0] if not isDR else [1, 2]
    for mode in modes:
        losses, dprobs = hp.loss_fwd_bwd(probs, b["label"], b["domain_id"], b["ob_label"], mode, 0.01, B)
        if mode == 0:
            lo = O.loss_cls(outs[0], outs[1], lab, dom)
            assert_close(losses[0], lo, 3e-5, 0)
        elif mode == 1:
            lc, le = O.loss_cls(outs[0], outs[1], lab, dom), O.loss_dr_e(*outs, lab, dom)
            assert_close(losses[0], lc, 3e-5, 0)
            assert_close(losses[1], le, 3e-5, 1e-7)
            lo = lc + 0.01 * le
        else:
            lo = O.loss_dr_r(*outs, lab, dom, ob)
            assert_close(losses[2], lo, 3e-5, 1e-7)
        for v in Po.values():
            v.grad = None
        lo.backward(retain_graph=True)
        G, ids_all, rows_all = hp.backward(m.param_dict(), m.cfg, ctx, dprobs)
        for k, v in Po.items():
            if k == "item_emb_layer.emb_item.weight":
                uid, ug, nu = hp.segreduce(ids_all, rows_all, V)
                dense = hp.dense_table_grad(uid, ug, nu, V)
                assert_close(dense, v.grad, 1e-3, grad_tol(v.grad), k)
            else:
                assert_close(G[k], v.grad, 1e-3, grad_tol(v.grad), f"{k} mode {mode}")


# ------------------------------------------------------------------ eval forward vs oracle, ragged shapes
@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("B,L,C,isInC,isItC", [(1, 1, 2, False, False), (7, 13, 5, False, True), (3, 130, 2, False, True),
                                              (4, 6, 4, True, True), (9, 200, 2, False, True), (2, 37, 50, True, False)])
def test_forward_vs_oracle_shapes(B, L, C, isInC, isItC, prec):
    rng = np.random.default_rng(B * 1000 + L)
    V = 211
    Le = 2 * L if isInC else L
    P = make_params(41, V, D, Le, HID, B, isInC=isInC, isItC=isItC)
    b = random_batch(rng, B, L, C, V)
    col = {}
    outs = oracle_forward(P, b, isInC=isInC, isItC=isItC, ts1=0.3, ts2=0.3, isDR=False, collect=col)
    if isItC:
        pj = torch.softmax(O.mim_scores(col["enc1"], col["enc2"]), 0)
        if (pj - 0.3).abs().min() < 1e-4:
            pytest.skip("gate margin too small")
    m = build_model(P, V, L, B, isInC=isInC, isItC=isItC, ts1=0.3, ts2=0.3, precision=prec).eval()
    with torch.no_grad():
        p1, p2 = run_model(m, to_cuda(b))
    assert_close(p1.reshape(B, C), outs[0], 0, 3e-5)
    assert_close(p2.reshape(B, C), outs[1], 0, 3e-5)


# ------------------------------------------------------------------ a6: MIM with gates on, reference literal golden
def test_mim_peaked_golden_forward_backward():
    hp = _hp()
    z = load("mim_peaked.npz")
    PP = {k: v.cuda() for k, v in make_params(15, 4, D, 6, HID, 12, isInC=True).items()}
    a, b = T(z["a"]).cuda().contiguous(), T(z["b"]).cuda().contiguous()
    wgt = T(z["wgt"]).cuda()
    for name, other, out_key, gkey, ga_key in (("itc_d1", b, "itc_out", "gitc/", "ga"), ("inc_d1", a, "inc_out", "ginc/", "ga_inc")):
        mg = hp._mim_scores(a, other, 6, None)
        st = hp._mim_forward(PP, name, mg, other, 6, 0.2, 0, None, want_esum=True)
        assert 0 < int(st.n_active.item()) < 12
        assert_close(st.E, z[out_key][0, 6:], 0, 3e-6)
        assert_close(st.esum, z[out_key][0, 6:].sum(0), 0, 2e-5)
        # backward: dOut = wgt ; dE = sum_i wgt[i, n:] ; d_self = wgt[:, :n]
        G = {k: torch.zeros_like(v) for k, v in PP.items()}
        dE = wgt[:, 6:].sum(0).contiguous()
        d_other = torch.zeros_like(other) if name == "itc_d1" else wgt[:, :6].clone().contiguous()
        hp._mim_backward(PP, G, st, dE, d_other, 0, None)
        for k in ("trans_nn.weight", "trans_nn.bias", "trans_bs.weight", "trans_bs.bias"):
            assert_close(G[f"{name}.{k}"], z[gkey + k], 1e-4, grad_tol(z[gkey + k]), k)
        if name == "itc_d1":
            assert_close(d_other, z["gb"], 1e-4, grad_tol(z["gb"]))
        else:
            assert_close(d_other, z[ga_key], 1e-4, grad_tol(z[ga_key]))


# ------------------------------------------------------------------ a9: segmented reduction + Adam
@pytest.mark.parametrize("n,V,hot", [(1, 10, 0), (100, 7, 0), (5000, 1000, 3000), (70000, 50, 0), (20000, 30000, 9000)])
def test_segreduce_deterministic_sorted_and_sums(n, V, hot):
    hp = _hp()
    g = torch.Generator().manual_seed(n)
    ids = torch.randint(0, V, (n,), generator=g)
    ids[:hot] = V - 1                                    # the pad-row hot spot
    ids = ids[torch.randperm(n, generator=g)].cuda()
    rows = torch.randn(n, D, generator=g).cuda()
    uid, ug, nu = hp.segreduce(ids, rows, V)
    k = int(nu.item())
    want_ids = torch.unique(ids)                         # sorted
    assert k == want_ids.numel() and torch.equal(uid[:k], want_ids)
    ref = torch.zeros(V, D, dtype=torch.float64, device="cuda").index_add_(0, ids, rows.double())
    assert_close(ug[:k], ref[want_ids].float(), 1e-4, 1e-4)
    uid2, ug2, nu2 = hp.segreduce(ids, rows, V)
    assert torch.equal(ug[:k], ug2[:k]), "segmented reduction must be bit-deterministic"
    assert_close(ug[:k].double().sum(0), rows.double().sum(0), 1e-5, 1e-3)    # conservation


def test_adam_dense_and_lazy_rows_match_reference_adam():
    from amid_b200._abi import call
    hp = _hp()
    g = torch.Generator().manual_seed(3)
    V, steps = 40, 9
    p0 = torch.randn(V, D, generator=g)
    pr, m, v = p0.clone(), torch.zeros(V, D), torch.zeros(V, D)                     # oracle: dense Adam
    pd, md, vd = p0.clone().cuda(), torch.zeros(V, D).cuda(), torch.zeros(V, D).cuda()   # ours: dense kernel
    pl, ml, vl = p0.clone().cuda(), torch.zeros(V, D).cuda(), torch.zeros(V, D).cuda()   # ours: lazy rows
    last = torch.zeros(V, dtype=torch.int32).cuda()
    for step in range(1, steps + 1):
        touched = torch.unique(torch.randint(0, V, (5,), generator=g))
        if step in (4, 5):
            touched = torch.tensor([0], dtype=torch.long)       # long gaps for the other rows
        grad = torch.zeros(V, D)
        grad[touched] = torch.randn(len(touched), D, generator=g)
        O.adam_step(pr, grad, m, v, step, 5e-4)
        gd = grad.cuda()
        call("amid_adam_dense", hp._ptr(pd), hp._ptr(gd), hp._ptr(md), hp._ptr(vd), V * D, step, 5e-4, 0.9, 0.999, 1e-8,
             hp._stream())
        uid = touched.cuda()
        ug = grad[touched].cuda().contiguous()
        nu = torch.tensor([len(touched)], dtype=torch.int32).cuda()
        call("amid_adam_rows_lazy", hp._ptr(pl), hp._ptr(ml), hp._ptr(vl), hp._ptr(last), hp._ptr(uid), hp._ptr(ug),
             hp._ptr(nu), len(touched), step, 5e-4, 0.9, 0.999, 1e-8, hp._stream())
    call("amid_adam_rows_flush", hp._ptr(pl), hp._ptr(ml), hp._ptr(vl), hp._ptr(last), V, steps, 5e-4, 0.9, 0.999, 1e-8,
         hp._stream())
    assert_close(pd, pr, 0, 2e-6)
    assert_close(pl, pr, 0, 2e-6)
    assert_close(ml, m, 1e-4, 1e-8)
    assert_close(vl, v, 1e-4, 1e-11)
    assert int(last.min().item()) in (0, steps)


# ------------------------------------------------------------------ a9: losses, all modes, incl. clamp edge cases
def test_loss_modes_vs_oracle_including_saturation():
    hp = _hp()
    g = torch.Generator().manual_seed(11)
    B, C = 9, 4
    probs = torch.rand(3, 2, B, C, generator=g) * 0.98 + 0.01
    probs[0, 0, 0, 0] = 1.0          # saturated sigmoid: loss 100 on the wrong label, zero gradient
    probs[0, 1, 1, 1] = 0.0
    probs[0, 0, 2, 1] = 1.0          # p = 1 on a negative: loss clamps at 100, BCELoss backward gives 1/1e-12
    lab = torch.cat((torch.ones(B, 1), torch.zeros(B, C - 1)), 1)
    dom = torch.randint(0, 2, (B,), generator=g)
    ob = torch.randint(0, 2, (B,), generator=g)
    for mode, nh in ((0, 1), (0, 3), (1, 3), (2, 3)):
        pc = probs[:nh].clone().requires_grad_(True)
        parts = [pc[h, k] for h in range(nh) for k in range(2)]
        if mode == 0:
            lo = O.loss_cls(parts[0], parts[1], lab, dom)
        elif mode == 1:
            lo = O.loss_cls(parts[0], parts[1], lab, dom) + 0.01 * O.loss_dr_e(*parts, lab, dom)
        else:
            lo = O.loss_dr_r(*parts, lab, dom, ob)
        lo.backward()
        losses, dprobs = hp.loss_fwd_bwd(probs[:nh].contiguous().cuda(), lab.cuda(), dom.cuda(), ob.cuda(), mode, 0.01, B)
        tot = losses[0] + 0.01 * losses[1] if mode == 1 else (losses[2] if mode == 2 else losses[0])
        assert_close(tot, lo, 2e-5, 0, f"mode {mode}")
        assert_close(dprobs, pc.grad, 2e-4, 1e-7, f"mode {mode}")


# ------------------------------------------------------------------ size-independent properties at C3 size
def test_c3_size_properties():
    """BASELINE.json config 3 shapes (B=1024, L=200): gather checksum, idempotence, finite outputs."""
    hp = _hp()
    from amid_b200._abi import call
    g = torch.Generator(device="cuda").manual_seed(0)
    V, B, L = 894820, 1024, 200
    table = torch.randn(V, D, device="cuda", generator=g)
    ids = torch.randint(0, V, (B * (2 * L + 2),), device="cuda", generator=g)
    out = torch.empty(ids.numel(), D, device="cuda")
    call("amid_emb_gather_fwd", hp._ptr(table), V, hp._ptr(ids), ids.numel(), hp._ptr(out), hp._stream())
    assert torch.equal(out, table[ids])
    out2 = torch.empty_like(out)
    call("amid_emb_gather_fwd", hp._ptr(table), V, hp._ptr(ids), ids.numel(), hp._ptr(out2), hp._stream())
    assert torch.equal(out, out2)
    # linearity of the segmented reduction: reduce(a + b) == reduce(a) + reduce(b) up to fp32 rounding
    a, b = torch.randn(50000, D, device="cuda", generator=g), torch.randn(50000, D, device="cuda", generator=g)
    sid = ids[:50000]
    u1, g1, n1 = hp.segreduce(sid, a, V)
    u2, g2, n2 = hp.segreduce(sid, b, V)
    u3, g3, n3 = hp.segreduce(sid, a + b, V)
    k = int(n1.item())
    assert torch.equal(u1[:k], u3[:k]) and (u1[:k][1:] > u1[:k][:-1]).all()
    assert_close(g3[:k], g1[:k] + g2[:k], 1e-4, 1e-5)


# ------------------------------------------------------------------ checkpoint / resume
@pytest.mark.parametrize("isDR", [False, True])
def test_checkpoint_resume_is_bit_exact(isDR):
    """Stop after two steps, restore into a fresh trainer, continue: identical bits to the uninterrupted run
    (dropout on, lazy row-Adam with pending rows at the time of the checkpoint, both optimizers when isDR)."""
    import io
    from amid_b200.engine import Trainer
    B, L, C, V = 8, 12, 2, 301
    rng = np.random.default_rng(5)
    P = make_params(44, V, D, L, HID, B, isDR=isDR)
    batches = [to_cuda(random_batch(rng, B, L, C, V)) for _ in range(5)]
    phases = [1, 2, 1, 1, 2] if isDR else [1] * 5

    def fresh():
        torch.manual_seed(1234)
        return Trainer(build_model(P, V, L, B, ts2=0.3, isDR=isDR, drop_p=0.5).train(), lr=1e-3, lr2=0.5)

    a = fresh()
    for b, ph in zip(batches[:2], phases[:2]):
        a.step(b, phase=ph)
    buf = io.BytesIO()
    torch.save(a.checkpoint(), buf)                   # through serialisation, as a real resume would
    for b, ph in zip(batches[2:], phases[2:]):
        a.step(b, phase=ph)
    a.flush()
    r = fresh()
    buf.seek(0)
    r.load_checkpoint(torch.load(buf, weights_only=False))
    for b, ph in zip(batches[2:], phases[2:]):
        r.step(b, phase=ph)
    r.flush()
    for n in a.P:
        assert torch.equal(a.P[n], r.P[n]), n
    for sa, sr in zip(a.opt, r.opt):
        assert sa.step == sr.step and torch.equal(sa.m, sr.m) and torch.equal(sa.tv, sr.tv)
    with pytest.raises(Exception):
        bad = a.checkpoint()
        bad["world"] = 4
        r.load_checkpoint(bad)


# ------------------------------------------------------------------ opt-in fast driver loops
def test_fast_test_loop_matches_reference_bookkeeping():
    """evaluate.test() == test() of train_sr.py:31-128 on the three real C1 eval batches of the golden fixture:
    same tuple layout, metrics identical to the list bookkeeping applied to the same scores, loss = the mean of the
    per-batch BCE means."""
    from amid_b200 import evaluate
    from amid_b200.engine import Trainer
    z = load("c1_eval_rank.npz")
    P = make_params(12, int(z["V"]), D, 20, HID, 256)
    tr = Trainer(build_model(P, int(z["V"]), 20, 256, ts2=0.4).eval())
    keys = ("i_node", "neg_samples", "seq_d1", "seq_d2", "domain_id", "overlap_label")
    loader = []
    for i in range(3):
        hb = {k: T(z[f"b{i}_{k}"]).float() for k in keys}                      # the reference collate yields float32
        C = hb["neg_samples"].shape[1] + 1
        hb["label"] = torch.cat((torch.ones(256, 1), torch.zeros(256, C - 1)), 1)
        loader.append(hb)
    out = evaluate.test(tr, loader, overlap=True)
    assert len(out) == 2 + 6 * 7
    # same scores through the list bookkeeping
    devb = [tr.to_device(h) for h in loader]
    probs = [tr.scores(b) for b in devb]
    p1, p2 = torch.cat([p[0, 0] for p in probs]), torch.cat([p[0, 1] for p in probs])
    dom, ov = torch.cat([b["domain_id"] for b in devb]), torch.cat([b["overlap_label"] for b in devb])
    res = evaluate.evaluate_lists(p1, p2, dom, ov)
    flat = []
    for k in ("d1_ov", "d1_no", "d2_ov", "d2_no", "d1", "d2"):
        flat += list(res[k])
    assert tuple(flat) == out[2:]
    # loss: mean over batches of mean(BCE(p_d1) * (1 - dom) + BCE(p_d2) * dom)  (train_sr.py:205-212)
    want = np.mean([float(O.loss_cls(probs[i][0, 0].cpu(), probs[i][0, 1].cpu(), devb[i]["label"].cpu(),
                                     devb[i]["domain_id"].cpu())) for i in range(3)])
    assert abs(out[0] - want) < 1e-6 and out[0] == out[1]
    short = evaluate.test(tr, loader, overlap=False)
    assert len(short) == 2 + 2 * 7 and short[2:] == out[-14:]


def test_fast_train_epoch_equals_manual_steps():
    from amid_b200.engine import Trainer
    B, L, C, V = 8, 12, 2, 97
    rng = np.random.default_rng(8)
    P = make_params(51, V, D, L, HID, B)
    host = [random_batch(rng, B, L, C, V) for _ in range(4)]

    def fresh():
        torch.manual_seed(77)
        return Trainer(build_model(P, V, L, B, ts2=0.3, drop_p=0.5).train(), lr=1e-3)

    a, b = fresh(), fresh()
    mean = a.train_epoch([{k: v.float() if k != "label" else v for k, v in h.items()} for h in host])
    ls = [float(b.step(to_cuda(h))[0]) for h in host]
    a.flush(); b.flush()
    for n in a.P:
        assert torch.equal(a.P[n], b.P[n]), n
    assert abs(mean - float(np.mean(ls))) < 1e-6


# ------------------------------------------------------------------ flag matrix, training direction
@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("isInC,isItC,isDR", [(True, True, False), (True, False, False), (False, False, False),
                                              (True, True, True)])
def test_flag_matrix_train_grads_vs_oracle(isInC, isItC, isDR, prec):
    """InnerComp in front of the encoders (model_seq.py:399-402, 422-424: positional table and attention over 2L),
    with / without InterComp, with / without the DR heads: probabilities, loss and every gradient (dropout off)."""
    hp = _hp()
    B, L, C, V = 6, 10, 3, 89
    Le = 2 * L if isInC else L
    rng = np.random.default_rng(17 + 2 * isInC + isItC)
    P = make_params(61, V, D, Le, HID, B, isInC=isInC, isItC=isItC, isDR=isDR)
    ts = 0.12
    m = build_model(P, V, L, B, isInC=isInC, isItC=isItC, ts1=ts, ts2=ts, isDR=isDR, drop_p=0.0, precision=prec).train()
    b = to_cuda(random_batch(rng, B, L, C, V))
    probs, ctx = hp.forward(m.param_dict(), m.cfg, b["i_node"], b["neg_samples"], b["seq_d1"], b["seq_d2"], train=True, seed=1)
    Po = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    col = {}
    outs = oracle_forward(Po, b, isInC=isInC, isItC=isItC, ts1=ts, ts2=ts, isDR=isDR, collect=col)
    if isItC:
        pj = torch.softmax(O.mim_scores(col["enc1"], col["enc2"]), 0)
        if (pj - ts).abs().min() < 1e-4:
            pytest.skip("gate margin too small")
    for i, o in enumerate(outs):
        assert_close(probs[i // 2, i % 2], o, 0, 3e-5, f"out {i}")
    mode = 1 if isDR else 0
    losses, dprobs = hp.loss_fwd_bwd(probs, b["label"], b["domain_id"], b["ob_label"], mode, 0.01, B)
    lab, dom = b["label"].cpu(), b["domain_id"].cpu()
    lo = O.loss_cls(outs[0], outs[1], lab, dom)
    assert_close(losses[0], lo, 3e-5, 0)
    if isDR:
        lo = lo + 0.01 * O.loss_dr_e(*outs, lab, dom)
    lo.backward()
    G, ids_all, rows_all = hp.backward(m.param_dict(), m.cfg, ctx, dprobs)
    for k, v in Po.items():
        if v.grad is None:
            continue
        if k == "item_emb_layer.emb_item.weight":
            uid, ug, nu = hp.segreduce(ids_all, rows_all, V)
            assert_close(hp.dense_table_grad(uid, ug, nu, V), v.grad, 1e-3, grad_tol(v.grad), k)
        else:
            assert_close(G[k], v.grad, 1e-3, grad_tol(v.grad), k)


@pytest.mark.parametrize("prec", PRECS)
def test_inc_itc_dr_training_direction_reference_golden(prec):
    """The same fixture that pins the oracle (tests/golden/make_inc_train_golden.py): InnerComp + InterComp + DR heads
    in training direction, executed by the reference -- outputs, phase-1 losses and gradients through the C ABI."""
    hp = _hp()
    z = load("inc_train_small.npz")
    V, ts, B, L = int(z["V"]), float(z["ts"]), 8, 6
    P = make_params(19, V, D, 2 * L, HID, B, isInC=True, isItC=True, isDR=True)
    m = build_model(P, V, L, B, isInC=True, isItC=True, ts1=ts, ts2=ts, isDR=True, drop_p=0.0, precision=prec).train()
    b = batch_from(z)
    probs, ctx = hp.forward(m.param_dict(), m.cfg, b["i_node"], b["neg_samples"], b["seq_d1"], b["seq_d2"], train=True, seed=3)
    for i, n in enumerate(("p1", "p2", "ips1", "ips2", "g1", "g2")):
        assert_close(probs[i // 2, i % 2], z[n], 0, 3e-5, n)
    losses, dprobs = hp.loss_fwd_bwd(probs, b["label"], b["domain_id"], b["ob_label"], 1, 0.01, B)
    assert_close(losses[0], z["loss_cls"], 3e-5, 0)
    assert_close(losses[1], z["loss_dr_e"], 3e-5, 1e-7)
    G, ids_all, rows_all = hp.backward(m.param_dict(), m.cfg, ctx, dprobs)
    n_checked = 0
    for k in z:
        if k.startswith("grad/"):
            assert_close(G[k.split("/", 1)[1]], z[k], 1e-3, grad_tol(z[k]), k)
            n_checked += 1
    assert n_checked == int(z["n_grad_tensors"])
    uid, ug, nu = hp.segreduce(ids_all, rows_all, V)
    want = np.zeros((V, D), dtype=np.float32)
    want[z["gtab_idx"]] = z["gtab_rows"]
    assert_close(hp.dense_table_grad(uid, ug, nu, V), want, 1e-3, grad_tol(want), "table")


@pytest.mark.parametrize("prec", PRECS)
def test_dr_phase2_reference_golden(prec):
    """The doubly-robust PHASE-2 step (loss_dr_r, train_sr_dr.py:392-394) against a fixture executed by the reference with
    dropout off (tests/golden/make_dr_phase2_golden.py): six outputs, the loss and the gradients through the C ABI."""
    hp = _hp()
    z = load("dr_phase2_nodrop.npz")
    V, ts, B, L = int(z["V"]), float(z["ts"]), 16, 20
    P = make_params(23, V, D, L, HID, B, isDR=True)
    m = build_model(P, V, L, B, ts2=ts, isDR=True, drop_p=0.0, precision=prec).train()
    b = batch_from(z)
    probs, ctx = hp.forward(m.param_dict(), m.cfg, b["i_node"], b["neg_samples"], b["seq_d1"], b["seq_d2"], train=True, seed=5)
    for i, n in enumerate(("p1", "p2", "ips1", "ips2", "g1", "g2")):
        assert_close(probs[i // 2, i % 2], z[n], 0, 3e-5, n)
    losses, dprobs = hp.loss_fwd_bwd(probs, b["label"], b["domain_id"], b["ob_label"], 2, 0.01, B)
    assert_close(losses[2], z["loss_dr_r"], 3e-5, 1e-7)
    G, ids_all, rows_all = hp.backward(m.param_dict(), m.cfg, ctx, dprobs)
    n_checked = 0
    for k in z:
        if k.startswith("grad/"):
            assert_close(G[k.split("/", 1)[1]], z[k], 1e-3, grad_tol(z[k]), k)
            n_checked += 1
    assert n_checked == int(z["n_grad_tensors"]) and n_checked >= 20
    uid, ug, nu = hp.segreduce(ids_all, rows_all, V)
    want = np.zeros((V, D), dtype=np.float32)
    want[z["gtab_idx"]] = z["gtab_rows"]
    assert_close(hp.dense_table_grad(uid, ug, nu, V), want, 1e-3, grad_tol(want), "table")

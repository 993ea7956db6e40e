"""tcgen05 / TMEM tile pipeline (amid_b200/csrc/tc.cuh) against torch fp32 references.
TF32 operands (10-bit mantissa), fp32 accumulation: tolerance 2e-3 relative to the row scale."""
import pytest
import torch

pytestmark = pytest.mark.gpu
D = 128


@pytest.mark.parametrize("M", [128, 1, 200, 1000, 4096 + 77])
def test_tc_linear_tf32(M):
    from amid_b200 import hotpath as hp
    from amid_b200._abi import call
    g = torch.Generator().manual_seed(M)
    x = torch.randn(M, D, generator=g).cuda()
    w = (torch.randn(D, D, generator=g) / 11.3).cuda()
    b = torch.randn(D, generator=g).cuda()
    y = torch.full((M, D), float("nan"), device="cuda")
    call("amid_tc_linear_test", hp._ptr(x), hp._ptr(w), hp._ptr(b), M, hp._ptr(y), hp._stream())
    torch.cuda.synchronize()
    ref = (x.double() @ w.double().T + b.double()).float()
    err = (y - ref).abs().max().item()
    assert torch.isfinite(y).all()
    assert err < 2e-3 * ref.abs().max().item(), err


@pytest.mark.parametrize("M", [128, 200, 4096 + 77])
def test_tc_linear_bf16(M):
    from amid_b200 import hotpath as hp
    from amid_b200._abi import call
    g = torch.Generator().manual_seed(M)
    x = torch.randn(M, D, generator=g).cuda()
    w = (torch.randn(D, D, generator=g) / 11.3).cuda()
    b = torch.randn(D, generator=g).cuda()
    y = torch.full((M, D), float("nan"), device="cuda")
    call("amid_tc_linear16_test", hp._ptr(x), hp._ptr(w), hp._ptr(b), M, hp._ptr(y), hp._stream())
    torch.cuda.synchronize()
    ref = (x.bfloat16().double() @ w.bfloat16().double().T + b.double()).float()     # bf16-rounded operands, exact accumulate
    assert (y - ref).abs().max().item() < 1e-4 * ref.abs().max().item() + 1e-4
    full = (x.double() @ w.double().T + b.double()).float()
    assert (y - full).abs().max().item() < 2e-2 * full.abs().max().item()


@pytest.mark.parametrize("M,ctas", [(128, 1), (1000, 3), (5000, 148)])
def test_tc_wgrad_bf16_mn_major(M, ctas):
    from amid_b200 import hotpath as hp
    from amid_b200._abi import call
    g = torch.Generator().manual_seed(M)
    dy = torch.randn(M, D, generator=g).cuda()
    x = torch.randn(M, D, generator=g).cuda()
    part = torch.full((ctas, D, D), float("nan"), device="cuda")
    call("amid_tc_wgrad16_test", hp._ptr(dy), hp._ptr(x), M, hp._ptr(part), ctas, hp._stream())
    torch.cuda.synchronize()
    ref = (dy.bfloat16().double().T @ x.bfloat16().double()).float()
    got = part.sum(0)
    assert (got - ref).abs().max().item() < 1e-4 * ref.abs().max().item() + 1e-3


@pytest.mark.parametrize("B,n", [(1, 1), (5, 7), (3, 64), (4, 200), (2, 130), (3, 256), (2, 300), (300, 200)])
def test_mim_scores_3xtf32_matches_fp32(B, n):
    from amid_b200 import hotpath as hp
    g = torch.Generator().manual_seed(B * 100 + n)
    a = torch.randn(B, n, D, generator=g).cuda()
    b = torch.randn(B, n, D, generator=g).cuda()
    m_tc = hp._mim_scores(a, b, n, None, True)
    m_ref = hp._mim_scores(a, b, n, None, False)
    want = torch.einsum("jsd,jtd->jst", a.double(), b.double()).flatten(1).max(1)[0].float()
    assert (m_ref - want).abs().max().item() < 1e-4
    assert (m_tc - want).abs().max().item() < 1e-4


# ------------------------------------------------------------------ the encoder on the tensor-core path
# TF32 operands: 10-bit mantissa, truncated by the MMA unit.  Stated tolerance for this path:
# probabilities 5e-3 abs, losses 3e-3 rel, gradient tensors 4e-2 in relative Frobenius norm (a ReLU
# pre-activation within TF32 noise of zero flips its gate and moves one term of a row of a weight
# gradient, so single elements are only held to 0.2 of the tensor's max magnitude).
import numpy as np


# per precision: (probabilities abs, loss rel, gradient relative-Frobenius, gradient element / max|g|, features abs)
TOL = {"tf32": (5e-3, 3e-3, 4e-2, 0.2, 3e-2), "bf16": (2e-2, 1e-2, 1.2e-1, 0.5, 1.2e-1)}
PRECISIONS = ["tf32", "bf16"]


def assert_grad_tc(got, ref, name, precision):
    got = got.detach().float().cpu()
    ref = torch.as_tensor(np.asarray(ref) if not torch.is_tensor(ref) else ref).float().cpu()
    num, den = (got - ref).norm().item(), ref.norm().item()
    assert num <= TOL[precision][2] * den + 1e-9, f"{name}: relative Frobenius error {num / max(den, 1e-30):.3e}"
    assert (got - ref).abs().max().item() <= TOL[precision][3] * ref.abs().max().item() + 1e-9, name


from helpers import (HID, O, T, assert_close, batch_from, build_model, grad_tol, load, make_params, oracle_forward,
                     random_batch, run_model, to_cuda)


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("B,L,C", [(1, 1, 2), (7, 13, 5), (3, 130, 2), (9, 200, 2), (64, 200, 2)])
def test_forward_tc_vs_oracle(B, L, C, precision):
    rng = np.random.default_rng(B * 1000 + L)
    V = 211
    P = make_params(41, V, D, L, HID, B)
    b = random_batch(rng, B, L, C, V)
    col = {}
    outs = oracle_forward(P, b, isInC=False, isItC=True, ts1=0.3, ts2=0.3, isDR=False, collect=col)
    pj = torch.softmax(O.mim_scores(col["enc1"], col["enc2"]), 0)
    if (pj - 0.3).abs().min() < 5e-2:
        pytest.skip("gate margin too small for a reduced-precision comparison")
    m = build_model(P, V, L, B, ts1=0.3, ts2=0.3, precision=precision).eval()
    with torch.no_grad():
        p1, p2 = run_model(m, to_cuda(b))
    assert_close(p1.reshape(B, C), outs[0], 0, TOL[precision][0])
    assert_close(p2.reshape(B, C), outs[1], 0, TOL[precision][0])
    from amid_b200 import hotpath as hp
    cb = to_cuda(b)
    _, ctx = hp.forward(m.param_dict(), m.cfg, cb["i_node"], cb["neg_samples"], cb["seq_d1"], cb["seq_d2"], train=False)
    assert_close(ctx.encs[0].view(B, L, D), col["enc1"], 0, TOL[precision][4])


@pytest.mark.parametrize("precision", PRECISIONS)
def test_train_p0_tc_golden(precision):
    from amid_b200.engine import Trainer
    z = load("train_p0.npz")
    V = int(z["V"])
    P = make_params(18, V, D, 20, HID, 16)
    m = build_model(P, V, 20, 16, ts2=0.07, drop_p=0.0, precision=precision).train()
    b = batch_from(z, pre="b0_")
    p1, p2 = run_model(m, b)
    crit = torch.nn.BCELoss(reduction="none")
    dom = b["domain_id"]
    loss = torch.mean(crit(p1, b["label"]) * (1 - dom).unsqueeze(1) + crit(p2, b["label"]) * dom.unsqueeze(1))
    assert_close(loss, z["loss_step0"], TOL[precision][1], 0)
    loss.backward()
    named = dict(m.named_parameters())
    for k in z:
        if k.startswith("grad/"):
            assert_grad_tc(named[k[5:]].grad, z[k], k, precision)
    m2 = build_model(P, V, 20, 16, ts2=0.07, drop_p=0.0, precision=precision).train()
    tr = Trainer(m2, lr=5e-4)
    for step in range(3):
        losses = tr.step(batch_from(z, pre=f"b{step}_"))
        assert_close(losses[0], z[f"loss_step{step}"], TOL[precision][1], 0, f"step {step}")


@pytest.mark.parametrize("precision", PRECISIONS)
def test_train_dropout_tc_vs_oracle(precision):
    from amid_b200 import hotpath as hp
    B, L, C, V = 6, 40, 2, 97
    rng = np.random.default_rng(77)
    P = make_params(31, V, D, L, HID, B)
    m = build_model(P, V, L, B, ts2=0.2, precision=precision).train()
    b = to_cuda(random_batch(rng, B, L, C, V))
    probs, ctx = hp.forward(m.param_dict(), m.cfg, b["i_node"], b["neg_samples"], b["seq_d1"], b["seq_d2"], train=True, seed=99)
    masks = {s: {k: v.cpu() for k, v in d.items()} for s, d in hp.dropout_masks(m.cfg, B, L, 99, "cuda").items()}
    Po = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    col = {}
    outs = oracle_forward(Po, b, isInC=False, isItC=True, ts1=0.5, ts2=0.2, isDR=False, masks=masks, collect=col)
    pj = torch.softmax(O.mim_scores(col["enc1"], col["enc2"]), 0)
    if (pj - 0.2).abs().min() < 5e-2:
        pytest.skip("gate margin too small for a reduced-precision comparison")
    assert_close(probs[0, 0], outs[0], 0, TOL[precision][0])
    lo = O.loss_cls(outs[0], outs[1], b["label"].cpu(), b["domain_id"].cpu())
    lo.backward()
    losses, dprobs = hp.loss_fwd_bwd(probs, b["label"], b["domain_id"], None, 0, 0.0, B)
    assert_close(losses[0], lo, TOL[precision][1], 0)
    G, ids_all, rows_all = hp.backward(m.param_dict(), m.cfg, ctx, dprobs)
    for k, v in Po.items():
        if k != "item_emb_layer.emb_item.weight":
            assert_grad_tc(G[k], v.grad, k, precision)


# ------------------------------------------------------------------ split-operand ("x3") primitives: fp32-level accuracy
@pytest.mark.parametrize("M,scale", [(128, 1.0), (1, 1.0), (200, 1e-6), (1000, 300.0), (4096 + 77, 1.0)])
def test_x3_linear_fp16_pairs_tmem_operand(M, scale):
    from amid_b200 import hotpath as hp
    from amid_b200._abi import call
    g = torch.Generator().manual_seed(M)
    x = torch.randn(M, D, generator=g) * scale
    x[:, 5] *= 1e-4                      # columns far below the row maximum keep their relative accuracy
    if M > 3:
        x[3] = 0.0                       # an all-zero row
    x = x.cuda()
    w = (torch.randn(D, D, generator=g) / 11.3).cuda()
    b = (torch.randn(D, generator=g) * scale).cuda()
    y = torch.full((M, D), float("nan"), device="cuda")
    scratch = torch.empty(65536 + 64, dtype=torch.uint8, device="cuda")
    call("amid_x3_linear_test", hp._ptr(x), hp._ptr(w), hp._ptr(b), M, hp._ptr(y), hp._ptr(scratch), hp._stream())
    torch.cuda.synchronize()
    ref = x.double() @ w.double().T + b.double()
    assert torch.isfinite(y).all()
    err = (y.double() - ref).abs().max().item()
    fp32 = ((x @ w.T + b).double() - ref).abs().max().item()     # what an fp32 GEMM leaves
    assert err < 2e-6 * ref.abs().max().item(), (err, fp32)


@pytest.mark.parametrize("M,ctas", [(64, 1), (128, 1), (130, 2), (1000, 3), (5000, 24)])
def test_x3_wgrad_bf16_triples(M, ctas):
    from amid_b200 import hotpath as hp
    from amid_b200._abi import call
    g = torch.Generator().manual_seed(M)
    dy = (torch.randn(M, D, generator=g) * torch.logspace(-6, 2, M).view(M, 1)).cuda()     # rows over 8 decades
    x = torch.randn(M, D, generator=g).cuda()
    wpart = torch.full((ctas, D, D), float("nan"), device="cuda")
    bpart = torch.full((ctas, D), float("nan"), device="cuda")
    call("amid_x3_wgrad_test", hp._ptr(dy), hp._ptr(x), M, hp._ptr(wpart), hp._ptr(bpart), ctas, hp._stream())
    torch.cuda.synchronize()
    ref = dy.double().T @ x.double()
    got = wpart.double().sum(0)
    assert (got - ref).abs().max().item() < 2e-6 * ref.abs().max().item()
    refb = dy.double().sum(0)
    assert (bpart.double().sum(0) - refb).abs().max().item() < 2e-6 * refb.abs().max().item()

"""Full-catalogue evaluation (BASELINE config 5) on the GPU: bit-identity with the sampled-candidate scorer,
integer-exact rank counts, and agreement with the oracle (predictModule over the whole pool)."""
import numpy as np
import pytest
import torch

from helpers import D, HID, build_model, oracle_forward  # noqa: F401
from common import make_params

pytestmark = pytest.mark.gpu


def _setup(B=16, L=12, V=3000, n1=700, n2=900, seed=5):
    from amid_b200.engine import Trainer
    P = make_params(seed, V, D, L, HID, B)
    m = build_model(P, V, L, B, ts2=0.07).eval()
    tr = Trainer(m)
    g = torch.Generator().manual_seed(seed)
    perm = torch.randperm(V, generator=g)
    pool_d1, pool_d2 = perm[:n1].clone(), perm[n1:n1 + n2].clone()
    dom = torch.randint(0, 2, (B,), generator=g)
    i_node = torch.where(dom == 0, pool_d1[torch.randint(0, n1, (B,), generator=g)],
                         pool_d2[torch.randint(0, n2, (B,), generator=g)])
    batch = {"i_node": i_node, "seq_d1": torch.randint(0, V, (B, L), generator=g),
             "seq_d2": torch.randint(0, V, (B, L), generator=g), "domain_id": dom,
             "overlap_label": torch.randint(0, 2, (B,), generator=g)}
    return P, m, tr, batch, pool_d1, pool_d2


def test_catalogue_scores_bit_identical_to_sampled_scorer_and_counts_exact():
    from amid_b200 import evaluate
    P, m, tr, batch, pool_d1, pool_d2 = _setup()
    dev = {k: v.cuda() for k, v in batch.items()}
    cat = tr.catalogue(pool_d1, pool_d2)
    res = evaluate.full_catalogue_ranks(tr.P, tr.cfg, cat, dev)
    for dom, pool in ((0, pool_d1), (1, pool_d2)):
        rows, ranks_fix, ranks_nofix = res[dom]
        assert np.array_equal(rows, np.nonzero(batch["domain_id"].numpy() == dom)[0])
        # the sampled path: every user scored against [positive, whole pool] as "negatives"
        neg = pool.view(1, -1).expand(len(batch["i_node"]), -1).contiguous().cuda()
        probs = tr.scores({**dev, "neg_samples": neg})[0, dom].cpu().numpy()        # [B, 1 + I]
        for k, r in enumerate(rows):
            pos_col = int(np.nonzero(pool.numpy() == int(batch["i_node"][r]))[0][0])
            others = np.delete(probs[r, 1:], pos_col)
            s0 = probs[r, 0]
            assert s0 == probs[r, 1 + pos_col]                                        # same pair, same bits
            if np.count_nonzero(others == s0) == 0:
                assert ranks_nofix[k] == int(np.count_nonzero(others > s0))
            s1 = np.float32(s0) - np.float32(1e-7)
            if np.count_nonzero(others == s1) == 0:
                assert ranks_fix[k] == int(np.count_nonzero(others > s1))


def test_catalogue_score_rows_match_sampled_scorer_bitwise():
    from amid_b200._abi import call
    from amid_b200 import hotpath as hp
    P, m, tr, batch, pool_d1, pool_d2 = _setup(B=8, n1=300, n2=257)
    dev = {k: v.cuda() for k, v in batch.items()}
    cat = tr.catalogue(pool_d1, pool_d2)
    B = 8
    _, ctx = hp.forward(tr.P, tr.cfg, dev["i_node"], dev["i_node"].view(B, 1).clone(), dev["seq_d1"], dev["seq_d2"],
                        train=False, seed=0, need_ctx=True)
    A = torch.empty(B, 2, 32, device="cuda")
    call("amid_catalogue_user_proj", hp._ptr(ctx.us[0]), hp._ptr(ctx.us[1]), B, hp._ptr(tr.P["predictModule.fc.0.weight"]), 32,
         hp._ptr(A), hp._stream())
    for dom, pool in ((0, pool_d1), (1, pool_d2)):
        lo, hi = cat.ranges[dom]
        rows = torch.arange(B, dtype=torch.int32, device="cuda")
        pos_idx = torch.full((B,), lo, dtype=torch.int32, device="cuda")
        sc = torch.empty(B, hi - lo, device="cuda")
        sp = torch.empty(B, device="cuda")
        call("amid_catalogue_scores", hp._ptr(A), hp._ptr(rows), B, dom, hp._ptr(cat.Bc), lo, hi, hp._ptr(pos_idx),
             hp._ptr(tr.P["predictModule.fc.2.weight"]), hp._ptr(tr.P["predictModule.fc.2.bias"]), hp._ptr(sp), hp._ptr(sc),
             hp._stream())
        neg = pool.view(1, -1).expand(B, -1).contiguous().cuda()
        probs = tr.scores({**dev, "neg_samples": neg})[0, dom]
        assert torch.equal(sc, probs[:, 1:])
        assert torch.equal(sp, probs[:, 1])


def test_catalogue_ranks_against_oracle():
    from oracle import amid_oracle as O
    from amid_b200 import evaluate
    P, m, tr, batch, pool_d1, pool_d2 = _setup(B=16, n1=500, n2=640, seed=9)
    dev = {k: v.cuda() for k, v in batch.items()}
    cat = tr.catalogue(pool_d1, pool_d2)
    res = evaluate.full_catalogue_ranks(tr.P, tr.cfg, cat, dev)
    rows_o = O.full_catalogue_scores(P, batch["i_node"], batch["seq_d1"], batch["seq_d2"], batch["domain_id"], pool_d1,
                                     pool_d2, isInC=False, isItC=True, ts1=0.5, ts2=0.07)
    want_fix = O.full_catalogue_ranks(rows_o, 1e-7)
    want_nofix = O.full_catalogue_ranks(rows_o, 0.0)
    for dom in (0, 1):
        rows, rf, rn = res[dom]
        # identical unless two scores are within fp32 noise of each other: allow a one-place move on a few users
        assert np.abs(rf - want_fix[rows]).max() <= 2
        assert np.abs(rn - want_nofix[rows]).max() <= 2
        assert np.count_nonzero(rf != want_fix[rows]) <= max(1, len(rows) // 4)
    met = evaluate.evaluate_full_catalogue(tr.P, tr.cfg, cat, [dev])
    assert set(met) >= {"d1", "d2"}
    for k in ("d1", "d2"):
        assert len(met[k]) == 7 and 0.0 <= met[k][6] <= 1.0


def test_catalogue_saturated_ties_fall_back_to_numpy_rule():
    """With the output bias pushed far positive every sigmoid saturates to 1.0: all pool items tie with the positive
    and the rank is whatever numpy's argsort makes of an all-equal row (utils.py:297)."""
    from amid_b200 import evaluate
    P, m, tr, batch, pool_d1, pool_d2 = _setup(B=8, n1=130, n2=70, seed=3)
    with torch.no_grad():
        tr.P["predictModule.fc.2.bias"].fill_(60.0)
        tr.P["predictModule.fc.2.weight"].zero_()
    dev = {k: v.cuda() for k, v in batch.items()}
    cat = tr.catalogue(pool_d1, pool_d2)
    res = evaluate.full_catalogue_ranks(tr.P, tr.cfg, cat, dev)
    for dom, n in ((0, 130), (1, 70)):
        if dom not in res:
            continue
        rows, rf, rn = res[dom]
        row = np.ones(n, dtype=np.float32)
        assert np.all(rn == (-row).argsort().argsort()[0])
        row[0] = row[0] - 1e-7
        assert np.all(rf == (-row).argsort().argsort()[0])


@pytest.mark.parametrize("graph", [False, True])
def test_fast_full_catalogue_equals_per_batch_path(graph):
    """evaluate_full_catalogue_fast (graph-replayed forwards, one rank launch per domain, one read-back, ties resolved at
    the end) returns exactly the metrics of the per-batch path, on several batches with overlap lists."""
    from amid_b200 import evaluate
    P, m, tr, batch, pool_d1, pool_d2 = _setup(B=16, n1=300, n2=410, seed=21)
    cat = tr.catalogue(pool_d1, pool_d2)
    rng = np.random.default_rng(4)
    batches = []
    for k in range(3):
        b = {kk: v.clone() for kk, v in batch.items()}
        perm = torch.from_numpy(rng.permutation(16))
        b = {kk: v[perm].contiguous() for kk, v in b.items()}
        b["overlap_label"] = torch.from_numpy(rng.integers(0, 2, 16))
        batches.append({kk: v.cuda() for kk, v in b.items()})
    want = evaluate.evaluate_full_catalogue(tr.P, tr.cfg, cat, batches)
    got = evaluate.evaluate_full_catalogue_fast(tr.P, tr.cfg, cat, batches, graph=graph)
    assert set(got) == set(want)
    for k in want:
        assert got[k] == want[k], k


def test_fast_full_catalogue_saturated_ties():
    from amid_b200 import evaluate
    P, m, tr, batch, pool_d1, pool_d2 = _setup(B=8, n1=130, n2=70, seed=3)
    with torch.no_grad():
        tr.P["predictModule.fc.2.bias"].fill_(60.0)
        tr.P["predictModule.fc.2.weight"].zero_()
    dev = {k: v.cuda() for k, v in batch.items()}
    cat = tr.catalogue(pool_d1, pool_d2)
    want = evaluate.evaluate_full_catalogue(tr.P, tr.cfg, cat, [dev, dev])
    got = evaluate.evaluate_full_catalogue_fast(tr.P, tr.cfg, cat, [dev, dev])
    assert got == want


def test_graphed_eval_forward_equals_eager():
    """CUDA-graph replay of the eval forward (C1 shape) returns the eager probabilities bit for bit, batch after batch."""
    from amid_b200 import evaluate
    P, m, tr, batch, pool_d1, pool_d2 = _setup(B=16, n1=300, n2=410, seed=5)
    dev = {k: v.cuda() for k, v in batch.items()}
    B, L = dev["seq_d1"].shape
    rng = np.random.default_rng(0)
    C = 7
    gf = evaluate.GraphedForward(tr.P, tr.cfg, B, L, C)
    for it in range(3):
        neg = torch.from_numpy(rng.integers(0, 700, (B, C - 1))).cuda()
        b = {**dev, "neg_samples": neg, "seq_d1": torch.roll(dev["seq_d1"], it, 0).contiguous()}
        want = tr.scores(b)
        got = gf.run(b)
        assert torch.equal(got, want)

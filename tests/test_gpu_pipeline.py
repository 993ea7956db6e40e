"""Device-side batch construction (csrc/pipeline.cu, amid_b200/pipeline.py) against the fixture produced by executing
the reference's DualDomainSeqDataset + collate_fn_enhance (tests/golden/make_dataset_golden.py)."""
import numpy as np
import pytest
import torch

from common import load, make_params
from helpers import D, HID, build_model

pytestmark = pytest.mark.gpu


def _dataset(z):
    from amid_b200.pipeline import DeviceDataset, prepare_rows
    n = int(z["n_rows"])
    s1 = [z["in_seq_d1_vals"][z["in_seq_d1_offs"][i]:z["in_seq_d1_offs"][i + 1]].tolist() for i in range(n)]
    s2 = [z["in_seq_d2_vals"][z["in_seq_d2_offs"][i]:z["in_seq_d2_offs"][i + 1]].tolist() for i in range(n)]
    prep = prepare_rows(z["in_user"].tolist(), s1, s2, z["in_domain"].tolist())
    return DeviceDataset(prep, int(z["seq_len"]), int(z["long_length"]), int(z["pad_id"])), s1, s2


@pytest.mark.parametrize("tag", ["train", "eval"])
def test_replay_mode_is_bit_identical_to_the_reference_batch(tag):
    z = load("dataset_small.npz")
    ds, _, _ = _dataset(z)
    n = len(ds)
    order = torch.arange(n)
    b = ds.batch(order, negatives=torch.from_numpy(z[f"{tag}_neg_samples"]).long(), check=True)
    for k in ("user_node", "i_node", "seq_d1", "seq_d2", "long_tail_mask_d1", "long_tail_mask_d2", "domain_id",
              "overlap_label", "neg_samples", "label"):
        want = z[f"{tag}_{k}"]
        got = b[k].cpu().numpy()
        assert got.shape == want.shape, k
        assert np.array_equal(got.astype(np.float64), want.astype(np.float64)), k
        assert b[k].dtype == (torch.float32 if k == "label" else torch.int64)
    # any order / repeated rows
    perm = torch.tensor([5, 0, 5, 38, 17])
    p = ds.batch(perm, negatives=torch.from_numpy(z[f"{tag}_neg_samples"][perm.numpy()]).long())
    assert np.array_equal(p["seq_d2"].cpu().numpy(), z[f"{tag}_seq_d2"][perm.numpy()].astype(np.int64))


def test_device_sampler_respects_the_reference_constraints():
    z = load("dataset_small.npz")
    ds, s1, s2 = _dataset(z)
    n = len(ds)
    rows = torch.arange(n)
    pools = [set(z["train_pool_d1"].tolist()), set(z["train_pool_d2"].tolist())]
    K = 7
    a = ds.batch(rows, k=K, seed=123, check=True)
    a2 = ds.batch(rows, k=K, seed=123)
    c = ds.batch(rows, k=K, seed=124)
    assert torch.equal(a["neg_samples"], a2["neg_samples"])                 # reproducible in (seed, row)
    assert not torch.equal(a["neg_samples"], c["neg_samples"])
    neg = a["neg_samples"].cpu().numpy()
    for i in range(n):
        dom = int(z["in_domain"][i])
        own = set(s1[i] if dom == 0 else s2[i])
        assert len(set(neg[i].tolist())) == K                               # without replacement (random.sample)
        assert all(x in pools[dom] and x not in own for x in neg[i].tolist())
    assert a["label"].shape == (n, K + 1)
    # every admissible item is reachable and the draw is not grossly non-uniform: row 3 (own domain d2, 7 items)
    counts = {}
    reps = 600
    for s in range(reps):
        x = int(ds.batch(torch.tensor([2]), k=1, seed=1000 + s)["neg_samples"][0, 0])
        counts[x] = counts.get(x, 0) + 1
    admissible = pools[1] - set(s2[2])
    assert set(counts) <= admissible and len(counts) >= 0.9 * len(admissible)
    exp = reps / len(admissible)
    assert max(counts.values()) < 3.0 * exp
    # the k draws of a row are a pseudo-random subset, not an arithmetic progression in pool order (the signature of an
    # affine walk over the sorted pool): equal consecutive gaps must be the exception
    dom = int(z["in_domain"][2])
    order = {v: i for i, v in enumerate(sorted(pools[dom]))}
    P = len(order)
    ap = 0
    for s in range(200):
        x = ds.batch(torch.tensor([2]), k=3, seed=5000 + s)["neg_samples"][0].tolist()
        p0, p1, p2 = (order[v] for v in x)
        ap += ((p1 - p0) % P) == ((p2 - p1) % P)
    assert ap < 60, ap


def test_sampler_reports_an_exhausted_pool_and_bad_rows():
    from amid_b200.pipeline import DeviceDataset, prepare_rows
    prep = prepare_rows([1, 2], [[1, 2, 3], [3, 4]], [[7], [8, 9]], [0, 1])   # row 0 may draw {4}, row 1 {7}
    ds = DeviceDataset(prep, 4, 2, 0)
    b = ds.batch(torch.tensor([0, 1]), k=1, check=True)
    assert b["neg_samples"].flatten().tolist() == [4, 7]
    with pytest.raises(IndexError):
        ds.batch(torch.tensor([1]), k=2, check=True)
    with pytest.raises(IndexError):
        ds.batch(torch.tensor([2]), k=1, check=True)
    with pytest.raises(ValueError):
        prepare_rows([1], [[]], [[5]], [0])                                  # empty own-domain history


def test_epoch_iterator_feeds_the_trainer():
    from amid_b200.engine import Trainer
    z = load("dataset_small.npz")
    ds, _, _ = _dataset(z)
    V, L, B = 100, int(z["seq_len"]), 8
    P = make_params(3, V, D, L, HID, B)
    tr = Trainer(build_model(P, V, L, B, ts2=0.3).train(), lr=1e-3)
    seen = 0
    for b in ds.epoch(B, k=1, seed=5):
        assert b["seq_d1"].shape == (B, L)
        losses = tr.step(b)
        seen += 1
    assert seen == len(ds) // B and torch.isfinite(losses).all()
    first = [b["i_node"].clone() for b in ds.epoch(B, seed=5)]
    again = [b["i_node"].clone() for b in ds.epoch(B, seed=5)]
    other = [b["i_node"].clone() for b in ds.epoch(B, seed=6)]
    assert all(torch.equal(x, y) for x, y in zip(first, again))
    assert any(not torch.equal(x, y) for x, y in zip(first, other))


def test_dr_variant_replay_matches_reference_batch():
    """DualDomainSeqDatasetDR + collate_fn_enhanceDR (dataset_seq.py:443-591): the same fields plus ob_label."""
    from amid_b200.pipeline import DeviceDataset, prepare_rows
    z = load("dataset_small.npz")
    n = int(z["n_rows"])
    s1 = [z["in_seq_d1_vals"][z["in_seq_d1_offs"][i]:z["in_seq_d1_offs"][i + 1]].tolist() for i in range(n)]
    s2 = [z["in_seq_d2_vals"][z["in_seq_d2_offs"][i]:z["in_seq_d2_offs"][i + 1]].tolist() for i in range(n)]
    prep = prepare_rows(z["in_user"].tolist(), s1, s2, z["in_domain"].tolist(), z["in_ob_label"].tolist())
    ds = DeviceDataset(prep, int(z["seq_len"]), int(z["long_length"]), int(z["pad_id"]))
    b = ds.batch(torch.arange(n), negatives=torch.from_numpy(z["dr_neg_samples"]).long(), check=True)
    for k in ("user_node", "i_node", "seq_d1", "seq_d2", "long_tail_mask_d1", "long_tail_mask_d2", "domain_id",
              "overlap_label", "ob_label", "neg_samples", "label"):
        assert np.array_equal(b[k].cpu().numpy().astype(np.float64), z[f"dr_{k}"].astype(np.float64)), k

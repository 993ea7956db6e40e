"""Sample construction (dataset_seq.py:177-236): the oracle restatement against the fixture produced by executing the
reference dataset + collate, and the host-side CSR preparation of amid_b200.pipeline against the oracle."""
import numpy as np

from common import load
from oracle import amid_oracle as O


def _rows(z):
    out = []
    for i in range(int(z["n_rows"])):
        s1 = z["in_seq_d1_vals"][z["in_seq_d1_offs"][i]:z["in_seq_d1_offs"][i + 1]].tolist()
        s2 = z["in_seq_d2_vals"][z["in_seq_d2_offs"][i]:z["in_seq_d2_offs"][i + 1]].tolist()
        out.append((int(z["in_user"][i]), s1, s2, int(z["in_domain"][i])))
    return out


def test_oracle_sample_construction_matches_reference_dataset():
    z = load("dataset_small.npz")
    L, ll, pad = int(z["seq_len"]), int(z["long_length"]), int(z["pad_id"])
    for tag in ("train", "eval"):
        pools = [set(z[f"{tag}_pool_d1"].tolist()), set(z[f"{tag}_pool_d2"].tolist())]
        for i, (u, s1, s2, dom) in enumerate(_rows(z)):
            s = O.build_sample(s1, s2, dom, L, ll, pad)
            assert s["i_node"] == z[f"{tag}_i_node"][i]
            assert s["seq_d1"] == z[f"{tag}_seq_d1"][i].tolist() and s["seq_d2"] == z[f"{tag}_seq_d2"][i].tolist()
            assert s["long_tail_mask_d1"] == z[f"{tag}_long_tail_mask_d1"][i]
            assert s["long_tail_mask_d2"] == z[f"{tag}_long_tail_mask_d2"][i]
            assert s["domain_id"] == z[f"{tag}_domain_id"][i] and s["overlap_label"] == z[f"{tag}_overlap_label"][i]
            # the reference's negatives come from pool - own sequence, without replacement
            negs = z[f"{tag}_neg_samples"][i].astype(np.int64).tolist()
            assert len(set(negs)) == len(negs)
            assert all(n in pools[dom] and n not in s["exclude"] for n in negs)
        assert z[f"{tag}_label"].shape[1] == z[f"{tag}_neg_samples"].shape[1] + 1
        assert (z[f"{tag}_label"][:, 0] == 1).all() and (z[f"{tag}_label"][:, 1:] == 0).all()


def test_host_csr_preparation_matches_oracle():
    from amid_b200.pipeline import prepare_rows
    z = load("dataset_small.npz")
    rows = _rows(z)
    prep = prepare_rows([r[0] for r in rows], [r[1] for r in rows], [r[2] for r in rows], [r[3] for r in rows])
    assert sorted(prep["pool_d1"].tolist()) == z["train_pool_d1"].tolist()
    assert sorted(prep["pool_d2"].tolist()) == z["train_pool_d2"].tolist()
    for i, (u, s1, s2, dom) in enumerate(rows):
        s = O.build_sample(s1, s2, dom, 10**6, 1, -1)       # no truncation: the whole processed histories
        for k, name in ((1, "seq_d1"), (2, "seq_d2")):
            got = prep[f"hist_d{k}_vals"][prep[f"hist_d{k}_offs"][i]:prep[f"hist_d{k}_offs"][i + 1]].tolist()
            assert got == [x for x in s[name] if x != -1]
        ex = prep["excl_vals"][prep["excl_offs"][i]:prep["excl_offs"][i + 1]].tolist()
        assert ex == sorted(s["exclude"])
        assert prep["target"][i] == s["i_node"] and prep["domain"][i] == s["domain_id"]
        assert prep["overlap"][i] == s["overlap_label"]


def test_dr_dataset_variant_only_adds_ob_label():
    """DualDomainSeqDatasetDR (dataset_seq.py:443-591) builds the same sample plus the row's ob_label."""
    from amid_b200.pipeline import prepare_rows
    z = load("dataset_small.npz")
    L, ll, pad = int(z["seq_len"]), int(z["long_length"]), int(z["pad_id"])
    rows = _rows(z)
    for i, (u, s1, s2, dom) in enumerate(rows):
        s = O.build_sample(s1, s2, dom, L, ll, pad)
        assert s["i_node"] == z["dr_i_node"][i] and s["seq_d1"] == z["dr_seq_d1"][i].tolist()
        assert s["seq_d2"] == z["dr_seq_d2"][i].tolist() and s["overlap_label"] == z["dr_overlap_label"][i]
    assert np.array_equal(z["dr_ob_label"].astype(np.int64), z["in_ob_label"])
    prep = prepare_rows([r[0] for r in rows], [r[1] for r in rows], [r[2] for r in rows], [r[3] for r in rows],
                        z["in_ob_label"].tolist())
    assert np.array_equal(prep["ob_label"], z["in_ob_label"])


def test_prepare_csv_reads_the_reference_layout(tmp_path):
    """The reference's CSV columns (dataset_seq.py:140-146, DR variant :446-453): JSON-encoded histories."""
    import json
    import pandas as pd
    from amid_b200.pipeline import prepare_csv, prepare_rows
    rows = [(7, [1, 2, 3], [50], 0, 1), (8, [], [51, 52], 1, 0), (9, [3, 3, 4], [], 0, 1)]
    p = tmp_path / "toy.csv"
    pd.DataFrame({"user_id": [r[0] for r in rows], "seq_d1": [json.dumps(r[1]) for r in rows],
                  "seq_d2": [json.dumps(r[2]) for r in rows], "domain_id": [r[3] for r in rows],
                  "ob_label": [r[4] for r in rows]}).to_csv(p, index=False)
    a = prepare_csv(str(p))
    b = prepare_rows([r[0] for r in rows], [r[1] for r in rows], [r[2] for r in rows], [r[3] for r in rows], [r[4] for r in rows])
    assert set(a) == set(b)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert a["target"].tolist() == [3, 52, 4]
    assert a["hist_d1_vals"].tolist() == [1, 2] + [] + [3, 3]          # row 2: target 4 removed, earlier 3s stay
    assert a["overlap"].tolist() == [1, 0, 0]


def test_device_dataset_refuses_cpu():
    import pytest
    from amid_b200 import _abi
    from amid_b200.pipeline import DeviceDataset, prepare_rows
    prep = prepare_rows([1], [[1, 2]], [[5]], [0])
    with pytest.raises(_abi.AmidError):
        DeviceDataset(prep, 4, 2, 0, device="cpu")

"""Golden fixture for the batch-construction path: EXECUTES the reference's DualDomainSeqDataset.__getitem__ +
collate_fn_enhance (dataset_seq.py:137-274) on a small synthetic CSV that exercises the edge cases of the sample
construction (target repeated earlier in its history, single-item histories, empty other-domain history, histories
shorter / equal / longer than seq_len).  Run once in the build container; tests only read the .npz.

Shim: random.sample on a set -> tuple (Python >= 3.11).
"""
import json
import os
import random
import sys

import numpy as np
import pandas as pd
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")

_orig_sample = random.sample


def _sample(pop, k, **kw):
    if isinstance(pop, (set, frozenset)):
        pop = tuple(pop)
    return _orig_sample(pop, k, **kw)


random.sample = _sample
import dataset_seq  # noqa: E402

dataset_seq.random.sample = _sample

SEQ_LEN, LONG_LEN, PAD = 6, 4, 99


def rows():
    rng = np.random.default_rng(0)
    d1_items, d2_items = list(range(0, 40)), list(range(50, 95))
    out = []

    def seq(pool, n):
        return [int(x) for x in rng.choice(pool, size=n, replace=True)]

    # hand-made edge cases first
    out.append((1, [3], [60, 61], 0))                              # own history of one item -> all pads after removal
    out.append((2, [5, 7, 5, 9, 5], [], 0))                        # target repeated earlier, other domain empty
    out.append((3, [], [70, 71, 72, 73, 74, 75, 76], 1))           # own = d2, exactly seq_len after removing the target
    out.append((4, [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11], [80], 0))  # longer than seq_len: keep the last seq_len
    out.append((5, [12, 13], [81, 82, 83, 84, 85, 86, 87, 88], 1)) # seq_len + 1 before removal
    out.append((6, [20, 21, 22, 23, 24, 25, 26], [90, 91], 0))     # exactly seq_len after removing the target
    for u in range(7, 40):
        dom = int(rng.integers(0, 2))
        n1, n2 = int(rng.integers(0, 12)), int(rng.integers(0, 12))
        if dom == 0 and n1 == 0:
            n1 = 1
        if dom == 1 and n2 == 0:
            n2 = 1
        out.append((u, seq(d1_items, n1), seq(d2_items, n2), dom))
    return out


def main():
    data = rows()
    csv = "/tmp/amid_dataset_golden.csv"
    pd.DataFrame({"user_id": [r[0] for r in data], "seq_d1": [json.dumps(r[1]) for r in data],
                  "seq_d2": [json.dumps(r[2]) for r in data], "domain_id": [r[3] for r in data]}).to_csv(csv, index=False)
    save = {"seq_len": SEQ_LEN, "long_length": LONG_LEN, "pad_id": PAD, "n_rows": len(data),
            "in_user": np.array([r[0] for r in data]), "in_domain": np.array([r[3] for r in data])}
    # ragged inputs as CSR
    for name, col in (("in_seq_d1", 1), ("in_seq_d2", 2)):
        save[name + "_vals"] = np.array([x for r in data for x in r[col]], dtype=np.int64)
        save[name + "_offs"] = np.cumsum([0] + [len(r[col]) for r in data]).astype(np.int64)
    for tag, is_train, neg in (("train", True, 1), ("eval", False, 5)):
        random.seed(11)
        ds = dataset_seq.DualDomainSeqDataset(SEQ_LEN, is_train, neg, LONG_LEN, PAD, csv)
        batch = dataset_seq.collate_fn_enhance([ds[i] for i in range(len(data))])
        for k, v in batch.items():
            assert v.dtype == torch.float32                    # the reference collate makes everything float32
            save[f"{tag}_{k}"] = v.numpy()
        save[f"{tag}_pool_d1"] = np.array(sorted(ds.item_pool_d1), dtype=np.int64)
        save[f"{tag}_pool_d2"] = np.array(sorted(ds.item_pool_d2), dtype=np.int64)
    # the doubly-robust variant (DualDomainSeqDatasetDR, dataset_seq.py:443-591): one more column, ob_label
    rng = np.random.default_rng(3)
    ob = rng.integers(0, 2, len(data))
    csv_dr = "/tmp/amid_dataset_golden_dr.csv"
    pd.DataFrame({"user_id": [r[0] for r in data], "seq_d1": [json.dumps(r[1]) for r in data],
                  "seq_d2": [json.dumps(r[2]) for r in data], "domain_id": [r[3] for r in data], "ob_label": ob}).to_csv(csv_dr, index=False)
    save["in_ob_label"] = ob.astype(np.int64)
    random.seed(12)
    ds = dataset_seq.DualDomainSeqDatasetDR(SEQ_LEN, True, 1, LONG_LEN, PAD, csv_dr)
    batch = dataset_seq.collate_fn_enhanceDR([ds[i] for i in range(len(data))])
    for k, v in batch.items():
        save[f"dr_{k}"] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "dataset_small.npz"), **save)
    print("wrote dataset_small.npz", {k: np.asarray(v).shape for k, v in save.items()})


if __name__ == "__main__":
    main()

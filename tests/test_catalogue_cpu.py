"""CPU checks of the full-catalogue oracle (oracle/amid_oracle.py, BASELINE config 5): scoring the whole pool must
reduce to the reference's sampled-candidate scoring when the pool is the sampled candidate set."""
import numpy as np
import torch

from common import make_params
from oracle import amid_oracle as O

D, HID = 128, 32


def test_full_catalogue_oracle_reduces_to_sampled_scoring():
    B, L, V = 6, 8, 400
    P = make_params(21, V, D, L, HID, B)
    g = torch.Generator().manual_seed(2)
    pool1, pool2 = torch.arange(0, 150), torch.arange(150, 400)
    dom = torch.tensor([0, 1, 0, 1, 1, 0])
    i_node = torch.where(dom == 0, pool1[torch.randint(0, 150, (B,), generator=g)], pool2[torch.randint(0, 250, (B,), generator=g)])
    s1, s2 = torch.randint(0, V, (B, L), generator=g), torch.randint(0, V, (B, L), generator=g)
    rows = O.full_catalogue_scores(P, i_node, s1, s2, dom, pool1, pool2, isInC=False, isItC=True, ts1=0.5, ts2=0.07)
    assert [len(r) for r in rows] == [150 if d == 0 else 250 for d in dom.tolist()]
    # sampled path of the reference: the same users with 5 of those pool items as negatives
    for r in range(B):
        pool = pool1 if dom[r] == 0 else pool2
        others = pool[pool != i_node[r]]
        neg = others[:5].view(1, 5).expand(B, 5)
        p = O.sasrec_forward(P, i_node, neg, s1, s2, isInC=False, isItC=True, ts1=0.5, ts2=0.07, isDR=False)[int(dom[r])]
        np.testing.assert_allclose(p[r].detach().numpy(), rows[r][:6], rtol=0, atol=2e-7)
    ranks = O.full_catalogue_ranks(rows, 1e-7)
    assert ranks.shape == (B,) and ranks.min() >= 0 and all(ranks[r] < len(rows[r]) for r in range(B))
    # the rank of the positive is the number of strictly better candidates when there is no tie
    for r in range(B):
        s = rows[r].copy()
        s[0] -= np.float32(1e-7)
        if np.count_nonzero(s[1:] == s[0]) == 0:
            assert ranks[r] == np.count_nonzero(s[1:] > s[0])

"""Shared helpers of the GPU parity tests (oracle = checker only, never the product path)."""
from __future__ import annotations

import numpy as np
import torch

from common import load, make_params  # noqa: F401  (tests/golden on sys.path via conftest)
from oracle import amid_oracle as O

D, HID = 128, 32


def T(a):
    return torch.from_numpy(np.asarray(a))


def build_model(P, V, L, bs, *, isInC=False, isItC=True, ts1=0.5, ts2=0.4, isDR=False, hid=HID, drop_p=0.5,
                precision="fp32"):
    """The drop-in SASRec on cuda:0 with the given reference-named parameters."""
    from amid_b200.model_seq import SASRec
    m = SASRec(user_length=10, user_emb_dim=D, item_length=V, item_emb_dim=D, seq_len=L, hid_dim=hid, bs=bs,
               isInC=isInC, isItC=isItC, threshold1=ts1, threshold2=ts2, isDR=isDR)
    res = m.load_state_dict(P, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    m.cfg.drop_p = drop_p
    m.cfg.precision = precision
    return m.cuda()


def batch_from(z, pre="in_", dev="cuda"):
    keys = ("i_node", "neg_samples", "seq_d1", "seq_d2", "domain_id", "label", "ob_label", "overlap_label")
    out = {}
    for k in keys:
        if pre + k in z:
            t = T(z[pre + k])
            out[k] = (t.float() if k == "label" else t.long()).to(dev).contiguous()
    return out


def run_model(m, b):
    B = b["seq_d1"].shape[0]
    dummy = torch.zeros(B, dtype=torch.long, device=b["seq_d1"].device)
    return m(dummy, b["i_node"], b["neg_samples"], b["seq_d1"], b["seq_d2"], dummy, dummy)


def oracle_forward(P, b, **kw):
    c = {k: v.cpu() for k, v in b.items()}
    return O.sasrec_forward(P, c["i_node"], c["neg_samples"], c["seq_d1"], c["seq_d2"], **kw)


def random_batch(rng, B, L, C, V, pad_frac=0.5):
    """Synthetic batch with left-padded histories (pad id = V-1) like dataset_seq.seq_padding."""
    seqs = []
    for _ in range(2):
        s = rng.integers(0, V - 1, size=(B, L))
        lens = rng.integers(1, L + 1, size=B)
        for i in range(B):
            if rng.random() < pad_frac:
                s[i, :L - lens[i]] = V - 1
        seqs.append(torch.from_numpy(s.astype(np.int64)))
    return {
        "seq_d1": seqs[0], "seq_d2": seqs[1],
        "i_node": torch.from_numpy(rng.integers(0, V - 1, size=B).astype(np.int64)),
        "neg_samples": torch.from_numpy(rng.integers(0, V - 1, size=(B, C - 1)).astype(np.int64)),
        "domain_id": torch.from_numpy(rng.integers(0, 2, size=B).astype(np.int64)),
        "ob_label": torch.from_numpy(rng.integers(0, 2, size=B).astype(np.int64)),
        "label": torch.cat((torch.ones(B, 1), torch.zeros(B, C - 1)), 1),
    }


def to_cuda(b):
    return {k: v.cuda().contiguous() for k, v in b.items()}


def assert_close(a, b, rtol, atol, msg=""):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, err_msg=msg)


def grad_tol(g, rel=2e-4):
    """absolute tolerance for a gradient tensor: rel * max|g| (fp32 accumulation-order noise)."""
    g = g.detach().cpu().numpy() if torch.is_tensor(g) else np.asarray(g)
    return 1e-8 + rel * float(np.abs(g).max())

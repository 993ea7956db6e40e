"""The vendored reference (oracle/_ref, see oracle/build_ref.py): it imports with the two shims, and the closed-form
InterComp patch used by the C3 reference arm reproduces the literal reference code."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref, ref_loader  # noqa: E402

needs_ref = pytest.mark.skipif(not ref_loader.available(), reason="oracle/_ref not built on this machine")


def test_recipe_reports_status():
    assert build_ref.build().startswith("oracle/_ref")


@needs_ref
def test_manifest_matches_files():
    import json
    man = json.load(open(os.path.join(build_ref.DST, "MANIFEST.json")))["files"]
    assert set(build_ref.CODE + build_ref.DATA) == set(man)
    for rel, dg in man.items():
        assert build_ref._sha(os.path.join(build_ref.DST, rel)) == dg


@needs_ref
def test_closed_form_itc_equals_literal_reference():
    ref = ref_loader.load("cpu")
    ms = ref.model_seq
    torch.manual_seed(3)
    B, n, d = 6, 5, 128
    lit = ms.InterComp(d, B, 0.12)
    a, b = torch.randn(B, n, d), torch.randn(B, n, d)
    with torch.no_grad():
        lit.trans_bs.weight.uniform_(-1, 1)
    want = lit(a, b)
    gates = torch.softmax(torch.einsum("jsd,jtd->jst", a, b).flatten(1).max(1)[0], 0) > 0.12
    assert 0 < int(gates.sum()) < B                      # some gates open, some closed
    literal_forward = ms.InterComp.forward
    try:
        ref_loader.closed_form_itc(ms)
        got = lit(a, b)
    finally:
        ms.InterComp.forward = literal_forward
    cuda_before = torch.Tensor.cuda
    with ref.on_cpu():
        assert torch.Tensor.cuda is not cuda_before
    assert torch.Tensor.cuda is cuda_before              # the CPU shim never leaks into the GPU arm
    assert torch.allclose(got, want, atol=2e-5, rtol=0), (got - want).abs().max()
    # gradients reach the same parameters
    (got.sum()).backward()
    assert lit.trans_nn.weight.grad is not None and lit.trans_bs.weight.grad is not None


@needs_ref
def test_reference_train_step_runs_on_cpu():
    ref = ref_loader.load("cpu")
    ms = ref.model_seq
    B, L, V = 4, 6, 50
    m = ms.SASRec(0, 128, V, 128, L, 32, B, False, True, 0.5, 0.4)
    opt = torch.optim.Adam(m.parameters(), lr=5e-4)
    ids = lambda *s: torch.randint(0, V, s)
    with ref.on_cpu():
        p1, p2 = m(ids(B), ids(B), ids(B, 1), ids(B, L), ids(B, L), ids(B), ids(B))
    lab = torch.cat((torch.ones(B, 1), torch.zeros(B, 1)), 1)
    loss = torch.nn.BCELoss()(p1, lab) + torch.nn.BCELoss()(p2, lab)
    opt.zero_grad(); loss.backward(); opt.step()
    assert torch.isfinite(loss)

"""CPU checks of the C-ABI library and the host-side mirror of the reference interface."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from common import load, make_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from amid_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "amid_b200.h")).read()
    declared = set(re.findall(r"\b(amid_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = ctypes.CDLL(built)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/amid_b200.h but not exported"
    from amid_b200 import _abi
    assert set(_abi.EXPORTS) == declared, set(_abi.EXPORTS) ^ declared
    assert _abi.lib().amid_version() == 100


def test_argument_validation_without_gpu(built):
    """Entry points reject bad arguments before touching the device (no compute call here)."""
    from amid_b200 import _abi
    l = _abi.lib()
    assert l.amid_emb_gather_fwd(None, 10, None, 4, None, None) < 0
    assert b"null" in l.amid_last_error()
    assert l.amid_rank_counts(None, 4, 3, 0.0, None, None, None) < 0
    with pytest.raises(_abi.AmidError):
        _abi.call("amid_score_fwd", None, None, None, None, 5, 32, 1, 2, None, None)
    assert l.amid_encoder_bwd_workspace_bytes(1024, 200) > 9 * 1024 * 200 * 128 * 4
    assert l.amid_embgrad_workspace_bytes(1000) > 0


def test_dropin_state_dict_names_shapes_and_signature():
    import inspect
    from amid_b200.model_seq import SASRec
    for isInC, isDR in ((False, False), (True, True)):
        m = SASRec(10, 128, 50, 128, 20, 32, 16, isInC, True, 0.5, 0.4, isDR)
        want = make_params(1, 50, 128, 40 if isInC else 20, 32, 16, isInC=isInC, isDR=isDR)
        got = m.state_dict()
        assert list(sorted(got)) == list(sorted(want))
        for k in want:
            assert tuple(got[k].shape) == tuple(want[k].shape), k
        assert all(p.is_leaf and p.requires_grad for p in m.parameters())
    sig = inspect.signature(SASRec.forward)
    assert list(sig.parameters)[1:] == ["u_node", "i_node", "neg_samples", "seq_d1", "seq_d2", "long_tail_mask_d1",
                                        "long_tail_mask_d2", "isTrain"]
    ctor = list(inspect.signature(SASRec.__init__).parameters)[1:]
    assert ctor == ["user_length", "user_emb_dim", "item_length", "item_emb_dim", "seq_len", "hid_dim", "bs", "isInC",
                    "isItC", "threshold1", "threshold2", "isDR"]


def test_dropin_refuses_cpu_and_wrong_width():
    from amid_b200 import AmidError
    from amid_b200.model_seq import SASRec
    with pytest.raises(AmidError):
        SASRec(10, 64, 50, 64, 20, 32, 16, False, True, 0.5, 0.4)
    m = SASRec(10, 128, 50, 128, 4, 32, 2, False, True, 0.5, 0.4)
    z = torch.zeros(2, dtype=torch.long)
    with pytest.raises(AmidError):            # CPU tensors: no fallback path
        m(z, z, torch.zeros(2, 1, dtype=torch.long), torch.zeros(2, 4, dtype=torch.long),
          torch.zeros(2, 4, dtype=torch.long), z, z)


def test_metrics_from_ranks_bit_exact_vs_reference_values():
    from amid_b200 import evaluate
    z = load("rank_ties.npz")
    assert evaluate.metrics_from_ranks(z["ranks"]) == tuple(z["met"].tolist())
    z = load("c1_eval_rank.npz")
    assert evaluate.metrics_from_ranks(z["ranks_d1"]) == tuple(z["met_d1"].tolist())
    assert evaluate.metrics_from_ranks(z["ranks_d2"]) == tuple(z["met_d2"].tolist())


def test_dropin_module_importable_as_model_seq():
    import importlib
    import sys
    sys.path.insert(0, os.path.join(ROOT, "amid_b200", "dropin"))
    try:
        sys.modules.pop("model_seq", None)
        ms = importlib.import_module("model_seq")
        assert ms.SASRec.__module__ == "amid_b200.model_seq"
    finally:
        sys.path.pop(0)
        sys.modules.pop("model_seq", None)


def test_bench_and_entry_points_compile_and_parse(monkeypatch):
    """bench.py / __graft_entry__.py are driver contracts: they must at least compile and parse their flags on a
    machine without a GPU (a module-level SyntaxError would only surface at round end otherwise)."""
    import importlib
    import py_compile
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for f in ("bench.py", "__graft_entry__.py", "tests/dp_check.py", "tests/eager_gpu_baseline.py"):
        py_compile.compile(os.path.join(root, f), doraise=True)
    sys.path.insert(0, root)
    bench = importlib.import_module("bench")
    monkeypatch.setattr(sys, "argv", ["bench.py", "--gpus", "2", "--steps", "5", "--warmup", "3", "--impl", "reference"])
    a = bench.parse()
    assert (a.gpus, a.steps, a.warmup, a.impl, a.precision, a.table_sync, a.ids) == (2, 5, 3, "reference", "x3", "auto", "uniform")
    w = bench.kernel_work("k_attn_bwd_mma", 1024, 200, 2)
    assert w["byte"] == 8 * 1024 * 200 * 128 * 4 and w["flop"] > 0
    rng = np.random.default_rng(0)
    b = bench.synth_batch(rng, 4, 6, 2, 100)
    assert b["seq_d1"].shape == (4, 6) and b["label"].shape == (4, 2)

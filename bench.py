"""bench.py -- training throughput of the AMID SASRec hot path on B200 (BASELINE.json metric
"train seqs/sec at 1/2/4/8 B200; emb-gather HBM GB/s; eval users/sec").

    python bench.py --gpus N --steps K --warmup W            # this repo (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference's own code on the host cores

A "step" is one full training step of train_sr.py:201-215 on one synthetic batch of the C3 shape (SURVEY.md
section 8): forward, domain-masked BCE, backward, embedding-gradient reduction and Adam on every parameter.
`value` = global sequences / second with the batch already resident in HBM; `e2e` = the same through the
public `Trainer` API with the batch in pinned HOST memory (H2D every step, loss read back every step).
The timed mode is the parity-grade one (`--precision x3`: tcgen05 tensor cores at fp32-level accuracy, the same
test tolerances as the exact fp32 path); `precisions` carries the other modes measured in the same run.
Prints ONE JSON line.

The reference arm drives the UNMODIFIED `model_seq.SASRec` of WujiangXu/AMID (vendored byte-for-byte into
oracle/_ref by oracle/build_ref.py, imported with the two SURVEY 8c shims) through the train_sr.py loop body with
torch.optim.Adam on the box's host cores.  At the C3 shape the literal InterComp needs 275 GB, so its forward is
swapped for the exact closed form (oracle/ref_loader.closed_form_itc; `kind` says so); each step runs a bounded
sample of the per-step batch so that the arm ends within minutes.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

V_ITEMS = 894820          # train_sr.py:450,456  item_length * 2
D, HID = 128, 32
REALISTIC = False
PRECISIONS = ("x3", "fp32", "tf32", "bf16")
PREC_NOTE = {
    "x3": "tcgen05 split-operand GEMMs + tcgen05 attention at fp32-level accuracy (parity-grade: the fp32 tolerances)",
    "fp32": "exact fp32 CUDA-core tiles (parity-grade anchor)",
    "tf32": "single-pass tcgen05 TF32 (reduced precision: 5e-3 tolerances)",
    "bf16": "single-pass tcgen05 BF16 operands (reduced precision: 2e-2 tolerances)",
}


def parse():
    global V_ITEMS, REALISTIC
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="per-GPU batch (weak scaling)")
    ap.add_argument("--seq-len", type=int, default=200)
    ap.add_argument("--neg", type=int, default=1, help="negatives per row in training (dataset_seq.py:197-199)")
    ap.add_argument("--dr", action="store_true", help="isDR=True, phase-1 loss (train_sr_dr.py config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the precisions / eval / c1 / dp_check sub-records")
    ap.add_argument("--cpu-sample", type=int, default=128, help="sequences per reference-arm step (bounded sample)")
    ap.add_argument("--cpu-budget", type=float, default=150.0, help="seconds the reference arm may spend on timed steps")
    ap.add_argument("--dense-table", action="store_true", help="dense Adam over the whole table (reference-style)")
    ap.add_argument("--table-sync", default="auto", choices=["auto", "sparse", "dense", "sharded"],
                    help="multi-GPU table strategy (engine.Trainer); 'sharded' = row-sharded table + all-to-all (config 4)")
    ap.add_argument("--ids", default="uniform", choices=["uniform", "realistic"],
                    help="uniform = roofline variant (default); realistic = short left-padded histories (SURVEY 8d)")
    ap.add_argument("--items", type=int, default=V_ITEMS, help="table rows V (config 4: 20000002)")
    ap.add_argument("--precision", default="x3", choices=list(PRECISIONS),
                    help="encoder arithmetic: x3 = tcgen05 at fp32-level accuracy (default, parity-grade), fp32 = exact "
                         "CUDA-core tiles, tf32 / bf16 = single-pass tcgen05 (reduced precision)")
    a = ap.parse_args()
    V_ITEMS = a.items
    REALISTIC = a.ids == "realistic"
    return a


def workload_name(a):
    return (f"C3 synthetic train step: per-GPU batch {a.batch}, L={a.seq_len}, d={D}, hid={HID}, C={1 + a.neg}, "
            f"V={V_ITEMS}, SASRec+ItC(ts2=0.4){'+DR' if a.dr else ''}, dropout 0.5, "
            f"{'uniform ids (roofline variant)' if a.ids == 'uniform' else 'realistic ids (median-5 histories, left-padded)'}")


def synth_batch(rng, B, L, C, V):
    """Uniform-random ids over [0,V), no padding (SURVEY.md section 8d roofline variant).  With --ids realistic the
    histories have min(L, Geometric) real items (median 5, as on cloth_sport) left-padded with the id V//2 + 1."""
    def seqs():
        s = rng.integers(0, V, (B, L))
        if REALISTIC:
            lens = np.minimum(L, rng.geometric(1.0 - 0.5 ** (1.0 / 5.0), B))
            s[np.arange(L)[None, :] < (L - lens)[:, None]] = V // 2 + 1
        return torch.from_numpy(s)
    return {
        "i_node": torch.from_numpy(rng.integers(0, V, B)),
        "neg_samples": torch.from_numpy(rng.integers(0, V, (B, C - 1))),
        "seq_d1": seqs(),
        "seq_d2": seqs(),
        "domain_id": torch.from_numpy(rng.integers(0, 2, B)),
        "ob_label": torch.from_numpy(rng.integers(0, 2, B)),
        "label": torch.cat((torch.ones(B, 1), torch.zeros(B, C - 1)), 1),
    }


# ------------------------------------------------------------------------------------------------
# the reference on the host cores
# ------------------------------------------------------------------------------------------------
def _ref():
    from oracle import ref_loader
    return ref_loader if ref_loader.available() else None


def ref_train_step(model, opt, crit, b, dr=False, dr_e_w=0.01):
    """Loop body of train_sr.py:191-215 (train_sr_dr.py:205-225 with --dr), device calls removed by the CPU shim."""
    outs = model(b["i_node"], b["i_node"], b["neg_samples"], b["seq_d1"], b["seq_d2"], b["i_node"], b["i_node"])
    dom = b["domain_id"]
    m1, m2 = (torch.ones_like(dom) - dom).unsqueeze(1), dom.unsqueeze(1)
    loss = torch.mean(crit(outs[0], b["label"]) * m1 + crit(outs[1], b["label"]) * m2)
    if dr:
        e1 = (crit(outs[0], b["label"]) - outs[4]) ** 2 / outs[2]
        e2 = (crit(outs[1], b["label"]) - outs[5]) ** 2 / outs[3]
        loss = loss + dr_e_w * torch.mean(e1 * m1 + e2 * m2)
    opt.zero_grad()
    loss.backward()
    opt.step()
    return float(loss.item())


def cpu_train_baseline(a, steps, warmup, sample, variant="closed_form_itc", budget=None):
    """Times train steps of the reference on `sample` sequences of the workload.  variant: 'closed_form_itc' (the
    configured SASRec+ItC with InterComp.forward swapped for its exact closed form) or 'no_itc' (isItC=False, the
    unpatched reference).  Falls back to the oracle port when oracle/_ref is absent.  Returns a cpu_baseline dict."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    L, C = a.seq_len, 1 + a.neg
    rng = np.random.default_rng(1234)
    rl = _ref()
    times = []
    if rl is not None:
        ref = rl.load("cpu")
        ms = ref.model_seq
        literal = ms.InterComp.forward
        if variant == "closed_form_itc":
            rl.closed_form_itc(ms)
        try:
          with ref.on_cpu():
            torch.manual_seed(7)
            model = ms.SASRec(0, D, V_ITEMS, D, L, HID, sample, False, variant != "no_itc", 0.5, 0.4, isDR=a.dr).train()
            opt = torch.optim.Adam(model.parameters(), lr=5e-4)          # train_sr.py:480
            crit = torch.nn.BCELoss(reduction="none")                    # train_sr.py:184
            t_start = time.perf_counter()
            for it in range(warmup + steps):
                b = synth_batch(rng, sample, L, C, V_ITEMS)
                t0 = time.perf_counter()
                ref_train_step(model, opt, crit, b, a.dr)
                dt = time.perf_counter() - t0
                if it >= warmup:
                    times.append(dt)
                    if budget is not None and len(times) >= 2 and time.perf_counter() - t_start + dt > budget:
                        break
        finally:
            ms.InterComp.forward = literal
        kind = "reference"
        what = ("the reference's model_seq.SASRec + train_sr.py loop body + torch.optim.Adam (dense 458 MB table), "
                + ("InterComp.forward swapped for its exact closed form (the literal [bs,bs,n,n] code needs 275 GB at this shape)"
                   if variant == "closed_form_itc" else "isItC=False (unpatched reference code)"))
    else:
        from common import make_keep_masks, make_params
        from oracle import amid_oracle as O
        P = {k: v.requires_grad_(True) for k, v in make_params(7, V_ITEMS, D, L, HID, sample, isDR=a.dr).items()}
        opt = torch.optim.Adam(list(P.values()), lr=5e-4)
        for it in range(warmup + steps):
            b = synth_batch(rng, sample, L, C, V_ITEMS)
            t0 = time.perf_counter()
            masks = make_keep_masks(it, sample, L, D)
            outs = O.sasrec_forward(P, b["i_node"], b["neg_samples"], b["seq_d1"], b["seq_d2"], isInC=False,
                                    isItC=variant != "no_itc", ts1=0.5, ts2=0.4, isDR=a.dr, masks=masks, closed_form=True)
            loss = O.loss_cls(outs[0], outs[1], b["label"], b["domain_id"])
            opt.zero_grad(); loss.backward(); opt.step()
            if it >= warmup:
                times.append(time.perf_counter() - t0)
        kind = "port"
        what = "oracle port of the reference path + torch.optim.Adam (oracle/_ref not built on this machine)"
    ms_step = 1e3 * float(np.mean(times))
    return {"value": sample / (ms_step / 1e3), "unit": "seq/s", "cores": cores, "kind": kind, "variant": variant,
            "ms_per_step": ms_step, "steps_timed": len(times),
            "sample": f"{sample} of the {a.batch} sequences of a per-GPU step (L={a.seq_len}), {warmup} warm-up + {len(times)} "
                      f"timed steps; {what}"}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_train_baseline(a, a.steps, min(a.warmup, 2), a.cpu_sample, "closed_form_itc", budget=a.cpu_budget)
    line = {
        "impl": "reference", "metric": "train_seqs_per_sec", "value": cb["value"], "unit": "seq/s", "n_gpus": a.gpus,
        "steps": cb["steps_timed"], "warmup": min(a.warmup, 2), "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "global_batch": a.batch * a.gpus, "seq_len": a.seq_len,
                   "parallelism": "host cores (the reference has no multi-GPU path: one process whatever --gpus says)"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "seq/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    try:
        line["no_itc"] = cpu_train_baseline(a, 2, 1, a.cpu_sample, "no_itc")
    except Exception as e:
        line["no_itc"] = {"value": None, "error": repr(e)}
    try:
        line["c1"] = c1_reference_cpu()
    except Exception as e:
        line["c1"] = {"value": None, "error": repr(e)}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# C1: real cloth_sport batches from the reference's own dataset + sampler + collate
# ------------------------------------------------------------------------------------------------
C1 = dict(bs=256, L=20, item_length=447410, neg_eval=999)


def c1_batches(n_train=6, n_eval=3):
    """Reference-collated batches (float32 tensors, train_sr.py:451-455 arguments) from the vendored CSVs, drawn ONCE
    (seeded) and fed to both arms: the same batches and the same sampled negatives."""
    import random
    rl = _ref()
    if rl is None:
        return None
    ref = rl.load("cpu")
    ds = ref.dataset_seq
    random.seed(1); np.random.seed(1); torch.manual_seed(1)
    L, bs = C1["L"], C1["bs"]
    pad_id = C1["item_length"] + 1                                       # train_sr.py:451
    root = ref.data_root

    def draw(csv, neg, n, is_train):
        d = ds.DualDomainSeqDataset(seq_len=L, isTrain=is_train, neg_nums=neg, long_length=7, pad_id=pad_id,
                                    csv_path=os.path.join(root, csv))
        out = []
        for k in range(n):
            out.append(ds.collate_fn_enhance([d[i] for i in range(k * bs, (k + 1) * bs)]))
        return out
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):          # the reference's dataset prints its sizes on stdout
        return {"train": draw("cloth_sport_train75.csv", C1["neg_eval"], n_train, True),
                "eval": draw("cloth_sport_test.csv", C1["neg_eval"], n_eval, False)}


def _c1_fields(b, long=True):
    f = (lambda t: t.long()) if long else (lambda t: t)
    return {"i_node": f(b["i_node"]), "neg_samples": f(b["neg_samples"]), "seq_d1": f(b["seq_d1"]), "seq_d2": f(b["seq_d2"]),
            "domain_id": f(b["domain_id"]), "label": b["label"].float()}


def c1_reference_cpu(batches=None):
    """The reference itself (LITERAL InterComp: it fits at bs=256, L=20) on the C1 batches, host cores."""
    rl = _ref()
    if rl is None:
        return {"value": None, "error": "oracle/_ref not built on this machine"}
    batches = batches or c1_batches()
    ref = rl.load("cpu")
    ms, ut = ref.model_seq, ref.utils
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    with ref.on_cpu():
        torch.manual_seed(5)
        model = ms.SASRec(0, D, C1["item_length"] * 2, D, C1["L"], HID, C1["bs"], False, True, 0.5, 0.4).train()
        opt = torch.optim.Adam(model.parameters(), lr=5e-4)
        crit = torch.nn.BCELoss(reduction="none")
        ts = []
        for k, hb in enumerate(batches["train"][:3]):
            b = _c1_fields(hb)
            t0 = time.perf_counter()
            ref_train_step(model, opt, crit, b)
            if k:
                ts.append(time.perf_counter() - t0)
        model.eval()
        te = []
        with torch.no_grad():
            for k, hb in enumerate(batches["eval"][:2]):
                b = _c1_fields(hb)
                t0 = time.perf_counter()
                p1, p2 = model(b["i_node"], b["i_node"], b["neg_samples"], b["seq_d1"], b["seq_d2"], b["i_node"], b["i_node"])
                pred = np.where((b["domain_id"] == 0).numpy()[:, None], p1.numpy(), p2.numpy())
                ut.get_sample_scores(pred)                                    # utils.py:296-313
                if k:
                    te.append(time.perf_counter() - t0)
    return {"kind": "reference", "cores": cores, "train_seqs_per_sec": C1["bs"] / float(np.mean(ts)),
            "eval_users_per_sec": C1["bs"] / float(np.mean(te)),
            "sample": f"cloth_sport_train75 / cloth_sport_test batches of {C1['bs']} rows, L={C1['L']}, eval C=1+{C1['neg_eval']}; "
                      f"unpatched reference (literal InterComp), {len(ts)} timed train steps, {len(te)} timed eval batches"}


def c1_ours(precision, graph_ok=True):
    """The same reference-drawn C1 batches through the public Trainer API from HOST tensors (H2D in the timed loop)."""
    from amid_b200 import evaluate
    from amid_b200.engine import Trainer
    from amid_b200.model_seq import SASRec
    batches = c1_batches()
    if batches is None:
        return {"value": None, "error": "oracle/_ref not built on this machine (C1 needs the reference's CSVs and sampler)"}
    torch.manual_seed(5)
    m = SASRec(0, D, C1["item_length"] * 2, D, C1["L"], HID, C1["bs"], False, True, 0.5, 0.4).cuda().train()
    m.cfg.precision = precision
    tr = Trainer(m, lr=5e-4)
    host = [{k: v.pin_memory() for k, v in _c1_fields(hb, long=False).items()} for hb in batches["train"]]
    for i in range(3):
        tr.step(tr.to_device(host[i % len(host)]))
    torch.cuda.synchronize()
    n = 60                                                                # one epoch of cloth_sport_train75 is 60 steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        losses = tr.step(tr.to_device(host[i % len(host)]))
    e1.record()
    torch.cuda.synchronize()
    train = n * C1["bs"] / (e0.elapsed_time(e1) / 1e3)
    m.eval()
    ev = [{k: v.pin_memory() for k, v in _c1_fields(hb, long=False).items()} for hb in batches["eval"]]

    def one(i):
        b = tr.to_device(ev[i % len(ev)])
        probs = tr.scores(b)
        return evaluate.evaluate_lists(probs[0, 0], probs[0, 1], b["domain_id"])
    for i in range(2):
        one(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(12):
        one(i)
    torch.cuda.synchronize()
    users = 12 * C1["bs"] / (time.perf_counter() - t0)
    return {"train_seqs_per_sec": train, "eval_users_per_sec": users, "precision": precision, "final_loss": float(losses[0].item()),
            "config": f"cloth_sport_train75 / cloth_sport_test batches from the reference's dataset + sampler + collate "
                      f"(float32 ids as collated), bs={C1['bs']}, L={C1['L']}, eval C=1+{C1['neg_eval']}, SASRec+ItC, 1 GPU, "
                      f"host batches through Trainer.step / Trainer.scores"}


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks DURING the timed region")
# ------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.3)                       # let the sampler reach its first sample before the timed region
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            out = ""
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm": p["hbm_gbs"], "tensor": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "src": "measured"}
    return {"hbm": 6650.0, "tensor": 1400.0, "src": "fallback"}


def kernel_work(name, B, L, C):
    """ALGORITHMIC work of ONE launch of a kernel at this workload: a dict with
    'flop' and/or 'byte'.  M = B*L tokens; act = one fp32 [M,128] activation tensor.  The byte figure counts
    every distinct activation tensor the kernel must read or write once (weights and statistics are noise)."""
    M = B * L
    act = M * D * 4.0
    gemm = 2.0 * M * D * D                      # one [M,128]x[128,128] contraction
    attn_full = 4.0 * B * L * L * D             # q k^T and P v over the full square, all 8 heads
    rows_seq, rows_items = M, B * C
    chain = {
        "k_ln_qkv": (3 * gemm, 5 * act),        # r: x            w: qn q k v
        "k_proj_ffn": (3 * gemm, 6 * act),      # r: o qn         w: x1 y h xout
        "k_ffn_bwd": (3 * gemm, 7 * act),       # r: dxo h x1     w: do2 dhpre dx1 dO
        "k_qkv_bwd": (3 * gemm, 6 * act),       # r: dq dk dv dx1 xin   w: dxin
        "k_wgrad": (6 * gemm, 11 * act),        # r: 6 dY + 5 distinct X
    }
    table = {
        "k_seq_embed": {"byte": rows_seq * (2 * D * 4 + 8)}, "k_gather": {"byte": rows_items * (2 * D * 4 + 8)},
        "k_embed_all": {"byte": (2 * rows_seq + rows_items) * (2 * D * 4 + 8)},
        "k_mim_scores": {"flop": 2.0 * B * L * L * D, "byte": 2 * act},
        "k_mim_scores_mma": {"flop": 2.0 * B * L * L * D, "byte": 2 * act},
        "k_mim_scores_tc5": {"flop": 3 * 2.0 * B * L * L * D, "byte": 2 * act},   # 3xTF32: three MMAs per product
    }
    for suffix in ("", "_mma", "_mma3", "_tc", "_p", "_t2"):
        table["k_attn_fwd" + suffix] = {"flop": attn_full, "byte": 4 * act}          # r: q k v     w: o
        table["k_attn_bwd" + suffix] = {"flop": 2.5 * attn_full, "byte": 8 * act}    # r: q k v o dO   w: dq dk dv
    for k, (f, b) in chain.items():
        for suffix in ("", "_tc", "_16", "_x3"):
            table[k + suffix] = {"flop": f, "byte": b}
    return table.get(name)


def measured_traffic(name, a):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed `ncu --set full` captures
    (profiles/r02_traffic_c3.json, else the round-1 file); only meaningful for the default C3 shape."""
    if (a.batch, a.seq_len, a.neg) != (1024, 200, 1):
        return None
    for f in ("r02_traffic_c3.json", "r01_traffic_c3.json"):
        path = os.path.join(ROOT, "profiles", f)
        if os.path.exists(path):
            v = json.load(open(path)).get("bytes_per_launch", {}).get(name)
            if v is not None:
                return v
    return None


def gather_gbs(model, a, iters=40):
    """BASELINE metric 2 ("emb-gather HBM GB/s"): amid_embed_all_fwd -- every table read of a train step (candidates +
    both histories, fused pos add / mask bits / dropout) -- launched back to back on rotating uniform-random id batches
    (SURVEY 8d roofline variant: the 458 MB table and the 210 MB of outputs per launch defeat L2 reuse), timed with
    CUDA events on the launch stream.  Bytes are ALGORITHMIC: R*(512 read + 512 write + 8 id), R = B*(2L+C)."""
    import ctypes as C_
    from amid_b200 import hotpath as hp
    from amid_b200._abi import call
    B, L, C = a.batch, a.seq_len, 1 + a.neg
    dev = torch.device("cuda")
    P = dict(model.named_parameters())
    table = P["item_emb_layer.emb_item.weight"].detach()
    pos1, pos2 = P["sac1.pos_emb.weight"].detach(), P["sac2.pos_emb.weight"].detach()
    g = torch.Generator(device="cuda").manual_seed(11)
    sets = [(torch.randint(0, V_ITEMS, (B, C), device=dev, generator=g), torch.randint(0, V_ITEMS, (B, L), device=dev, generator=g),
             torch.randint(0, V_ITEMS, (B, L), device=dev, generator=g)) for _ in range(4)]
    items = torch.empty(B, C, D, device=dev)
    x0 = [torch.empty(B * L, D, device=dev) for _ in range(2)]
    tm = [torch.empty(B * L * 4, device=dev, dtype=torch.int32) for _ in range(2)]
    s = hp._stream()

    def one(i):
        it, s1, s2 = sets[i % 4]
        drop = hp._dropout(model.cfg, True, 1000 + i, 0)
        call("amid_embed_all_fwd", hp._ptr(table), V_ITEMS, hp._ptr(it), B * C, hp._ptr(s1), hp._ptr(s2), hp._ptr(pos1),
             hp._ptr(pos2), B, L, hp._ptr(items), hp._ptr(x0[0]), hp._ptr(x0[1]), hp._ptr(tm[0]), hp._ptr(tm[1]),
             C_.byref(drop), s)

    for i in range(5):
        one(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        one(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    nbytes = B * (2 * L + C) * (2 * D * 4 + 8)
    return nbytes / (us * 1e-6) / 1e9, us


# ------------------------------------------------------------------------------------------------
# evaluation metrics (BASELINE metric 3 and config 5), each with its CPU baseline and roofline
# ------------------------------------------------------------------------------------------------
def eval_full_catalogue_users_per_sec(a, pk, dctx=None, world=1, rank=0, users_per_rank=125_000):
    """BASELINE config 5: every user ranked against the whole item pool of its target domain (16,084 / 12,153 items: the
    cloth_sport_train75 pools), 256-user eval batches (L=20, the C1 eval shape, the MIM couples the users of a
    batch), 125,000 users per GPU -- 1,000,000 users on 8 GPUs -- whole batches sharded across ranks, rank lists combined
    by one tensor all-gather.  Wall clock from the first batch to the final metrics (graph-replayed forwards, one rank
    launch per domain, one read-back), max over ranks.  Every rank calls this; rank 0 returns the record."""
    import torch.distributed as dist
    from amid_b200 import evaluate
    from amid_b200.engine import Trainer
    from amid_b200.model_seq import SASRec
    B, L = 256, 20
    n1, n2 = 16084, 12153
    torch.manual_seed(2)
    m = SASRec(0, D, V_ITEMS, D, L, HID, B, False, True, 0.5, 0.4, isDR=False).cuda().eval()
    m.cfg.precision = a.precision
    tr = Trainer(m)
    g = torch.Generator(device="cuda").manual_seed(9)
    perm = torch.randperm(V_ITEMS, device="cuda", generator=g)[:n1 + n2]
    pool1, pool2 = perm[:n1].contiguous(), perm[n1:].contiguous()
    cat = tr.catalogue(pool1, pool2)
    nb = (users_per_rank + B - 1) // B
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    ri = lambda hi, *s: torch.randint(0, hi, s, device="cuda", generator=g)
    batches = []
    for _ in range(nb):
        dom = ri(2, B)
        batches.append({"seq_d1": ri(V_ITEMS, B, L), "seq_d2": ri(V_ITEMS, B, L), "domain_id": dom,
                        "i_node": torch.where(dom == 0, pool1[ri(n1, B)], pool2[ri(n2, B)]), "overlap_label": ri(2, B)})

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    evaluate.evaluate_full_catalogue_fast(tr.P, tr.cfg, cat, batches[:4], dctx)         # warm-up (graph capture included)
    barrier()
    t0 = time.perf_counter()
    res = evaluate.evaluate_full_catalogue_fast(tr.P, tr.cfg, cat, batches, dctx)
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = float(dt.item())
    # the per-batch path of round 1 on a few batches, for reference
    t0 = time.perf_counter()
    evaluate.evaluate_full_catalogue(tr.P, tr.cfg, cat, batches[:12])
    torch.cuda.synchronize()
    dt_slow = (time.perf_counter() - t0) / 12
    if rank != 0:
        return None
    users = nb * B * world
    # the U x I stage alone at catalogue scale: 16,384 users of one domain against its pool (kernel time, CUDA events)
    from amid_b200 import hotpath as hp
    from amid_b200._abi import call
    nu = 16384
    A = torch.randn(nu, 2, 32, device="cuda")
    rows = torch.arange(nu, device="cuda", dtype=torch.int32)
    pos_idx = torch.randint(0, n1, (nu,), device="cuda", dtype=torch.int32)
    counts = torch.empty(nu, 4, device="cuda", dtype=torch.int32)
    s_pos = torch.empty(nu, device="cuda")
    w2, b2 = tr.P["predictModule.fc.2.weight"], tr.P["predictModule.fc.2.bias"]

    def rank_once():
        call("amid_catalogue_rank", hp._ptr(A), hp._ptr(rows), nu, 0, hp._ptr(cat.Bc), 0, n1, hp._ptr(pos_idx), hp._ptr(w2),
             hp._ptr(b2), 1e-7, hp._ptr(counts), hp._ptr(s_pos), hp._stream())

    rank_once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        rank_once()
    e1.record()
    torch.cuda.synchronize()
    kernel_pairs = 5.0 * nu * n1 / (e0.elapsed_time(e1) / 1e3)
    # the U x I stage is FP32-ALU bound (ReLU couples user and item inside the nonlinearity): ~165 issued
    # instructions per (user, item) pair against 148 SMs x 128 lanes at the SM clock
    alu_peak = 148 * 128 * 1.965e9 / 165.0
    out = {"metric": "eval_users_per_sec_full_catalogue", "value": users / dt, "unit": "users/s", "n_gpus": world,
           "users": users, "seconds": dt, "pairs_per_sec": users * (n1 + n2) / 2.0 / dt,
           "per_batch_path_users_per_sec": B / dt_slow,
           "kernel_pairs_per_sec": kernel_pairs, "kernel_users_per_sec": kernel_pairs / n1,
           "metrics_d1": list(res.get("d1", ())),
           "roofline": {"kernel": "k_rank_full", "bound": "fp32-alu", "achieved": kernel_pairs / 1e9, "peak": alu_peak / 1e9,
                        "unit": "Gpair/s", "frac": kernel_pairs / alu_peak, "traffic": None,
                        "note": "peak = 148 SMs x 128 fp32 lanes x 1.965 GHz / 165 instructions per pair"},
           "config": f"{users} users = {world} GPU(s) x {nb} batches of 256, L={L}, pools {n1} / {n2} items (every pool item scored, "
                     f"fp32 post-sigmoid), HR/NDCG/MRR of six lists; wall clock incl. forwards, ranking, read-back and gather"}
    if not a.no_cpu_baseline:
        try:
            from oracle import amid_oracle as O
            Pc = {k: v.detach().cpu() for k, v in tr.P.items()}
            hb = {k: v.cpu() for k, v in batches[0].items()}
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            t0 = time.perf_counter()
            with torch.no_grad():
                O.full_catalogue_scores(Pc, hb["i_node"], hb["seq_d1"], hb["seq_d2"], hb["domain_id"], pool1.cpu(), pool2.cpu(),
                                        isInC=False, isItC=True, ts1=0.5, ts2=0.4)
            dtc = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": B / dtc, "unit": "users/s", "cores": cores, "kind": "port",
                                   "sample": f"one 256-user batch against the whole pool of each user's domain: oracle forward + "
                                             f"predictModule over the pool (the reference itself never scores a full catalogue, "
                                             f"dataset_seq.py:201)"}
        except Exception as e:
            out["cpu_baseline"] = {"value": None, "kind": "port", "sample": f"failed: {e!r}"}
    return out


def eval_users_per_sec(a, pk, n_batches=20):
    """BASELINE metric 3 ("eval users/sec"): the C1 evaluation shape of run.sh -- 256 users per batch,
    1 + 999 candidates, L = 20 -- scored in eval mode and ranked on the device (test() of train_sr.py:31-128).
    One GPU (rank 0), batches resident in HBM, includes the D2H of the rank counts and the metric reduce."""
    from amid_b200 import evaluate
    from amid_b200.engine import Trainer
    from amid_b200.model_seq import SASRec
    B, L, C = 256, 20, 1000
    torch.manual_seed(1)
    m = SASRec(0, D, V_ITEMS, D, L, HID, B, False, True, 0.5, 0.4, isDR=False).cuda().eval()
    m.cfg.precision = a.precision
    tr = Trainer(m)
    rng = np.random.default_rng(7)
    hosts = [synth_batch(rng, B, L, C, V_ITEMS) for _ in range(4)]
    bs = [tr.to_device(h) for h in hosts]

    def one(i):
        b = bs[i % 4]
        probs = tr.scores(b)
        return evaluate.evaluate_lists(probs[0, 0], probs[0, 1], b["domain_id"])

    for i in range(3):
        one(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n_batches):
        one(i)
    torch.cuda.synchronize()
    eager = n_batches * B / (time.perf_counter() - t0)
    gf = evaluate.GraphedForward(tr.P, tr.cfg, B, L, C)                   # the launch-bound forward as one graph replay

    def one_g(i):
        b = bs[i % 4]
        probs = gf.run(b)
        return evaluate.evaluate_lists(probs[0, 0], probs[0, 1], b["domain_id"])

    for i in range(3):
        one_g(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n_batches):
        one_g(i)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    users = n_batches * B / dt
    bytes_user = (2 * L + C) * (2 * D * 4 + 8)                            # 1,073,280 B per user at C=1000
    out = {"metric": "eval_users_per_sec", "value": users, "unit": "users/s", "ms_per_batch": 1e3 * dt / n_batches,
           "eager_users_per_sec": eager,
           "roofline": {"kernel": "k_embed_all (candidate + history gather: the only O(C) HBM traffic of an eval batch)",
                        "bound": "hbm", "achieved": users * bytes_user / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                        "frac": users * bytes_user / 1e9 / pk["hbm"], "traffic": None,
                        "note": "end-to-end users/s x algorithmic gather bytes per user; a 256-user batch is bound by the "
                                "host-side ranking read-back, not by HBM"},
           "config": f"{B} users x {C} candidates per batch, L={L}, SASRec+ItC, CUDA-graph forward (evaluate.GraphedForward), "
                     f"device ranking + HR/NDCG/MRR, 1 GPU; eager_users_per_sec = the same without the graph"}
    rl = _ref()
    if not a.no_cpu_baseline and rl is not None:
        try:
            ref = rl.load("cpu")
            ms, ut = ref.model_seq, ref.utils
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            with ref.on_cpu():
                torch.manual_seed(1)
                rm = ms.SASRec(0, D, V_ITEMS, D, L, HID, B, False, True, 0.5, 0.4).eval()
                ts = []
                with torch.no_grad():
                    for k in range(2):
                        b = hosts[k]
                        t0 = time.perf_counter()
                        p1, p2 = rm(b["i_node"], b["i_node"], b["neg_samples"], b["seq_d1"], b["seq_d2"], b["i_node"], b["i_node"])
                        pred = np.where((b["domain_id"] == 0).numpy()[:, None], p1.numpy(), p2.numpy())
                        ut.get_sample_scores(pred)
                        if k:
                            ts.append(time.perf_counter() - t0)
            out["cpu_baseline"] = {"value": B / float(np.mean(ts)), "unit": "users/s", "cores": cores, "kind": "reference",
                                   "sample": "one timed 256-user x 1000-candidate batch (after one warm-up): unpatched reference "
                                             "model_seq.SASRec (literal InterComp) + utils.get_sample_scores"}
        except Exception as e:
            out["cpu_baseline"] = {"value": None, "kind": "reference", "sample": f"failed: {e!r}"}
    return out


# ------------------------------------------------------------------------------------------------
# data-parallel correctness record (rank 0 replays the global batch on one GPU)
# ------------------------------------------------------------------------------------------------
def dp_check(dctx, world, rank, precision):
    """Small-shape DP step (dropout ON, per-rank masks) against a single-GPU replay of the same global batch on rank 0:
    max relative differences of the loss, of the all-reduced dense gradients, and of the parameters after the step,
    for the table strategies that apply.  Every rank takes part in the collectives; rank 0 returns the record."""
    from common import make_params
    from amid_b200.engine import Trainer
    from amid_b200.model_seq import SASRec
    Bl, L, C, V = 8, 24, 2, 512
    Bg = Bl * world
    P = make_params(9, V, D, L, HID, Bg)
    rng = np.random.default_rng(3)
    b = synth_batch(rng, Bg, L, C, V)

    def build():
        m = SASRec(10, D, V, D, L, HID, Bg, False, True, 0.5, 0.3)
        m.load_state_dict(P)
        m.cfg.precision = precision
        return m.cuda().train()

    out = {}
    for mode in ("sparse", "dense", "sharded"):
        from amid_b200.hotpath import DistCtx
        tr = Trainer(build(), lr=1e-3, dist=DistCtx(), table_sync=mode)
        shard = {k: v[rank * Bl:(rank + 1) * Bl].cuda().contiguous() for k, v in b.items()}
        loss = tr.step(shard).clone()
        g = tr.flat_g.clone()
        tr.flush()
        table = tr.full_table().clone()
        params = tr.flat_p.clone()
        torch.cuda.synchronize()
        if rank == 0:
            ref = Trainer(build(), lr=1e-3)
            ref.model._seed_base = tr.model._seed_base
            l1 = ref.step({k: v.cuda().contiguous() for k, v in b.items()})
            ref.flush()
            rel = lambda x, y: float((x - y).norm() / y.norm().clamp_min(1e-30))
            out[mode] = {"loss_rel": abs(float(loss[0]) - float(l1[0])) / max(abs(float(l1[0])), 1e-30),
                         "dense_grad_rel": rel(g, ref.flat_g), "dense_param_rel": rel(params, ref.flat_p),
                         "table_rel": rel(table[:V], ref.table.data[:V])}
        del tr
    if rank == 0:
        worst = max(max(v.values()) for v in out.values())
        out["ok"] = bool(worst < 1e-4)
        out["config"] = (f"{world} ranks x {Bl} sequences, L={L}, V={V}, dropout 0.5 with row-indexed masks, one Adam step; "
                         f"relative Frobenius differences against a single-GPU replay of the global batch on rank 0")
    return out if rank == 0 else None


# ------------------------------------------------------------------------------------------------
# BASELINE config 4: 10M items per domain, row-sharded table + all-to-all lookups (multi-GPU only)
# ------------------------------------------------------------------------------------------------
def sharded_20m(a, dctx, world, rank, steps=6):
    """The C3 step with V = 20,000,002 rows (10.2 GB table): the table is never replicated -- rank r keeps rows
    {r, r+G, ...} and their Adam state, every step fetches the rows it reads with one device-planned all-to-all lookup
    and returns the gradient rows to their owners (amid_b200/sharded.py).  Every rank calls this; rank 0 returns."""
    import torch.distributed as dist
    from amid_b200.engine import Trainer
    from amid_b200.model_seq import SASRec
    V = 20_000_002
    B, L, C = a.batch, a.seq_len, 1 + a.neg
    torch.manual_seed(0)
    with torch.device("cuda"):
        model = SASRec(user_length=0, user_emb_dim=D, item_length=V, item_emb_dim=D, seq_len=L, hid_dim=HID, bs=B * world,
                       isInC=False, isItC=True, threshold1=0.5, threshold2=0.4, isDR=False)
    model = model.cuda().train()
    model.cfg.precision = a.precision
    from amid_b200.hotpath import DistCtx
    tr = Trainer(model, lr=5e-4, dist=DistCtx(), table_sync="sharded")    # its own context: the table routing hangs off it
    torch.cuda.empty_cache()
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    ri = lambda hi, *s: torch.randint(0, hi, s, device="cuda", generator=g)        # int64 ids made on the device
    bs = [{"i_node": ri(V, B), "neg_samples": ri(V, B, C - 1), "seq_d1": ri(V, B, L), "seq_d2": ri(V, B, L),
           "domain_id": ri(2, B), "label": torch.cat((torch.ones(B, 1), torch.zeros(B, C - 1)), 1).cuda()} for _ in range(3)]
    for i in range(2):
        tr.step(bs[i % 3])
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        losses = tr.step(bs[i % 3])
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item()) / steps
    mem = torch.cuda.max_memory_allocated() / 2**30
    loss = float(losses[0].item())
    del tr, model
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    return {"metric": "train_seqs_per_sec", "value": B * world / (ms / 1e3), "unit": "seq/s", "ms_per_step": ms, "steps": steps,
            "n_gpus": world, "precision": a.precision, "final_loss": loss, "peak_hbm_gib_rank0": mem,
            "config": f"V = {V} rows (10.2 GB fp32 table, never replicated), row-sharded over {world} GPUs, per-GPU batch {B}, "
                      f"L={L}, uniform int64 ids generated on the device, lookup plan on the device + NCCL all-to-all"}


# ------------------------------------------------------------------------------------------------
def run_ours(a):
    from amid_b200 import _abi
    from amid_b200.engine import Trainer
    from amid_b200.hotpath import DistCtx
    from amid_b200.model_seq import SASRec
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the amid_b200 hot path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dctx = None
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created; stdout must carry
        # only the JSON line, so fd 1 points at stderr until the first collective has completed.
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            warm = torch.ones(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
        dctx = DistCtx()
    B, L, C = a.batch, a.seq_len, 1 + a.neg
    Bg = B * world
    torch.manual_seed(0)
    model = SASRec(user_length=0, user_emb_dim=D, item_length=V_ITEMS, item_emb_dim=D, seq_len=L, hid_dim=HID, bs=Bg,
                   isInC=False, isItC=True, threshold1=0.5, threshold2=0.4, isDR=a.dr).cuda().train()
    model.cfg.precision = a.precision
    # distinct table rows a rank touches per step: every position for uniform ids, ~8 real items per history otherwise
    hint = B * (2 * L + C) if a.ids == "uniform" else B * (2 * 8 + C)
    tr = Trainer(model, lr=5e-4, dist=dctx, sparse_table=not a.dense_table, rows_per_step_hint=hint,
                 table_sync=a.table_sync)
    rng = np.random.default_rng(100 + rank)
    n_pool = 4
    host = [{k: v.pin_memory() for k, v in synth_batch(rng, B, L, C, V_ITEMS).items()} for _ in range(n_pool)]
    devb = [tr.to_device(h) for h in host]
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    dev_step = lambda i: tr.step(devb[i % n_pool])

    def host_step(i):
        losses = tr.step(tr.to_device(host[i % n_pool]))
        return losses.cpu()                          # D2H of the step's loss

    for i in range(a.warmup):
        dev_step(i)
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    l0 = _abi.kernel_launches()
    ms_total = timed(dev_step, a.steps)
    launches = _abi.kernel_launches() - l0
    for i in range(2):
        host_step(i)
    ms_e2e = timed(host_step, a.steps)
    clk = clocks.stop() if rank == 0 else None
    h2d = sum(v.numel() * (4 if k == "label" else 8) for k, v in host[0].items())

    # the other precision modes on the same trainer, same batches (all ranks take part: the steps contain collectives)
    precisions = {}
    if not a.no_extras:
        for mode in PRECISIONS:
            if mode == a.precision:
                precisions[mode] = {"seqs_per_sec": Bg / (ms_total / a.steps / 1e3), "ms_per_step": ms_total / a.steps,
                                    "steps": a.steps, "note": PREC_NOTE[mode] + " -- the timed mode of this line"}
                continue
            model.cfg.precision = mode
            n = 6 if mode == "fp32" else 10
            for i in range(2):
                dev_step(i)
            ms = timed(dev_step, n)
            precisions[mode] = {"seqs_per_sec": Bg / (ms / n / 1e3), "ms_per_step": ms / n, "steps": n, "note": PREC_NOTE[mode]}
        model.cfg.precision = a.precision

    # per-kernel CUDA-event profile of 3 more steps of the same workload (rank 0 reports).  The two encoder
    # chains are serialised for these steps: with both streams active a kernel's event interval would also
    # contain its neighbour's work.
    model.cfg.overlap_encoders = False
    _abi.profile(True)
    prof_steps = 3
    for i in range(prof_steps):
        dev_step(i)
    rep = _abi.profile_report()
    _abi.profile(False)
    model.cfg.overlap_encoders = True
    final_loss = float(tr.last_losses[0].item())

    dpc = None
    if world > 1 and not a.no_extras:
        try:
            dpc = dp_check(dctx, world, rank, a.precision)
        except Exception as e:
            dpc = {"ok": False, "error": repr(e)}
    sh20 = None
    if world > 1 and not a.no_extras and a.items == 894820:
        try:
            sh20 = sharded_20m(a, dctx, world, rank)
        except Exception as e:
            sh20 = {"value": None, "error": repr(e)}
    efc = None
    if not a.no_extras and tr.table_sync != "sharded":
        try:
            efc = eval_full_catalogue_users_per_sec(a, peaks(), dctx, world, rank)
        except Exception as e:
            efc = {"value": None, "error": repr(e)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    ms_step = ms_total / a.steps
    tot = sum(ms for _, ms in rep.values()) or 1.0
    breakdown = []
    for name, (cnt, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
        ent = {"kernel": name, "launches_per_step": cnt / prof_steps, "ms_per_step": ms / prof_steps,
               "share": ms / tot}
        w = kernel_work(name, B, L, C)
        if w:
            # the binding roofline is the one with the larger lower bound on the launch time
            per_launch_s = (ms / cnt) / 1e3
            t_hbm = w.get("byte", 0.0) / (pk["hbm"] * 1e9)
            t_tc = w.get("flop", 0.0) / (pk["tensor"] * 1e12)
            if t_tc > t_hbm:
                ent.update(bound="tensor", achieved=w["flop"] / per_launch_s / 1e12, unit="TFLOP/s")
                ent["frac"] = ent["achieved"] / pk["tensor"]
            else:
                ent.update(bound="hbm", achieved=w["byte"] / per_launch_s / 1e9, unit="GB/s")
                ent["frac"] = ent["achieved"] / pk["hbm"]
            if "flop" in w:
                ent["tflops"] = w["flop"] / per_launch_s / 1e12
        breakdown.append(ent)
    dom = next((e for e in breakdown if "bound" in e), None)
    roofline = None
    if dom:
        roofline = {"kernel": dom["kernel"], "bound": dom["bound"], "achieved": dom["achieved"],
                    "peak": pk["tensor"] if dom["bound"] == "tensor" else pk["hbm"], "unit": dom["unit"],
                    "frac": dom["frac"], "traffic": measured_traffic(dom["kernel"], a), "peak_source": pk["src"],
                    "share_of_step": dom["share"]}
    gat = next((e for e in breakdown if e["kernel"] == "k_embed_all"), None)
    roofline_gather = None
    if gat is not None:
        roofline_gather = {"kernel": "k_embed_all", "bound": "hbm", "achieved": gat["achieved"], "peak": pk["hbm"],
                           "unit": "GB/s", "frac": gat["frac"], "traffic": None, "peak_source": pk["src"],
                           "timing": "inside the train step (per-kernel events, serialised streams)"}
        try:
            if tr.table_sync == "sharded":
                raise RuntimeError("row-sharded table: the isolated gather runs on the replicated layout only")
            gbs, us = gather_gbs(model, a)
            roofline_gather.update(achieved=gbs, frac=gbs / pk["hbm"], us_per_launch=us, in_step_gbs=gat["achieved"],
                                   timing="back-to-back launches on rotating uniform-random id batches (CUDA events); "
                                          "in_step_gbs is the same kernel timed inside the train step")
        except Exception as e:                     # keep the in-step number
            roofline_gather["isolated_error"] = repr(e)
    # tensor-pipe utilisation is quoted on the 12*L*d^2 projection/FFN part only (fwd + 2x bwd),
    # over the time of the kernels that hold those contractions.  `issued` counts the MMAs the split-operand mode
    # actually issues for them (3 FP16-pair products per chain GEMM, 6 BF16-triple products per weight gradient).
    strip = lambda n: n.split("_tc")[0].split("_16")[0].split("_x3")[0]
    gemm_ms = sum(e["ms_per_step"] for e in breakdown
                  if strip(e["kernel"]) in ("k_ln_qkv", "k_proj_ffn", "k_ffn_bwd", "k_qkv_bwd", "k_wgrad"))
    gemm_flop = 3.0 * 4 * 12 * L * D * D * B          # 2 blocks x 2 encoders, fwd + dX + dW
    issued = {"x3": (2 * 3 + 6) / 3.0}.get(a.precision, 1.0)
    tensor_pipe = None if gemm_ms <= 0 else {
        "flop_per_step": gemm_flop, "ms_in_gemm_kernels": gemm_ms, "achieved_tflops": gemm_flop / (gemm_ms / 1e3) / 1e12,
        "peak_tflops": pk["tensor"], "frac": gemm_flop / (gemm_ms / 1e3) / 1e12 / pk["tensor"],
        "issued_mma_tflops": issued * gemm_flop / (gemm_ms / 1e3) / 1e12,
        "issued_frac": issued * gemm_flop / (gemm_ms / 1e3) / 1e12 / pk["tensor"],
        "note": "d=128 chains move 5-11 fp32 activation tensors per 3-6 GEMMs (32-64 flop/B): HBM-bound, see per-kernel frac; "
                "issued_* counts the extra piece products of the fp32-accurate split"}
    line = {
        "metric": "train_seqs_per_sec", "value": Bg / (ms_step / 1e3), "unit": "seq/s", "n_gpus": world,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"fp32": "f32", "x3": "f32", "tf32": "tf32", "bf16": "bf16"}[a.precision],
        "data": "synthetic",
        "config": {"workload": workload_name(a), "global_batch": Bg, "seq_len": L, "parallelism": f"dp{world}",
                   "precision": a.precision + ": " + PREC_NOTE[a.precision],
                   "l2": "no explicit flush: each step streams ~8 GB of activations, far larger than the 126 MB L2",
                   "table_update": ("dense" if a.dense_table else "row-sparse lazy Adam (exact dense semantics)") if world == 1
                   else f"table_sync={tr.table_sync}"},
        "clocks": clk,
        "e2e": {"value": Bg / (ms_e2e / a.steps / 1e3), "unit": "seq/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 12, "ms_per_step": ms_e2e / a.steps},
        "gpu_launches": launches,
        "roofline": roofline,
        "roofline_gather": roofline_gather,
        "tensor_pipe": tensor_pipe,
        "precisions": precisions,
        "kernel_breakdown": breakdown[:12],
        "kernel_tail_ms": {e["kernel"]: round(e["ms_per_step"], 4) for e in breakdown[12:]},
        "final_loss": final_loss,
    }
    if dpc is not None:
        line["dp_check"] = dpc
    if sh20 is not None:
        line["sharded_20M"] = sh20
    if not a.no_extras:
        try:
            line["eval"] = eval_users_per_sec(a, pk)
        except Exception as e:
            line["eval"] = {"value": None, "error": repr(e)}
        line["eval_full_catalogue"] = efc
        if world == 1:
            try:
                line["c1"] = c1_ours(a.precision)
                if not a.no_cpu_baseline:
                    line["c1"]["cpu_baseline"] = c1_reference_cpu()
            except Exception as e:
                line["c1"] = {"value": None, "error": repr(e)}
    if world == 1 and not a.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_train_baseline(a, 2, 1, a.cpu_sample, "closed_form_itc")
        except Exception as e:  # the baseline is reported beside the number, never a reason to lose it
            line["cpu_baseline"] = {"value": None, "unit": "seq/s", "cores": os.cpu_count(), "kind": "reference",
                                    "sample": f"failed: {e!r}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)

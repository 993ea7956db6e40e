"""bench.py -- training throughput of the AMID SASRec hot path on B200 (BASELINE.json metric
"train seqs/sec at 1/2/4/8 B200").

    python bench.py --gpus N --steps K --warmup W            # this repo (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference path's CPU port on the host cores

A "step" is one full training step of train_sr.py:201-215 on one synthetic batch of the C3
shape: forward, domain-masked BCE, backward, embedding-gradient
reduction and Adam on every parameter.  `value` = global sequences / second with the batch
already resident in HBM; `e2e` = the same through the public `Trainer` API with the batch in
pinned HOST memory (H2D every step, loss read back every step).  Prints ONE JSON line.
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
    ap.add_argument("--cpu-sample", type=int, default=32, help="sequences per CPU-baseline step (bounded sample)")
    ap.add_argument("--dense-table", action="store_true", help="dense Adam over the whole table (reference-style)")
    ap.add_argument("--table-sync", default="auto", choices=["auto", "sparse", "dense", "sharded"],
                    help="multi-GPU table strategy (engine.Trainer); 'sharded' = row-sharded table + all-to-all (config 4)")
    ap.add_argument("--ids", default="uniform", choices=["uniform", "realistic"],
                    help="uniform = roofline variant (default); realistic = short left-padded histories (SURVEY 8d)")
    ap.add_argument("--items", type=int, default=V_ITEMS, help="table rows V (config 4: 20000002)")
    ap.add_argument("--precision", default="bf16", choices=["fp32", "x3", "tf32", "bf16"],
                    help="encoder GEMM stages: tf32 = tcgen05 tensor cores (fp32 accumulate), fp32 = exact CUDA-core tiles")
    a = ap.parse_args()
    V_ITEMS = a.items
    REALISTIC = a.ids == "realistic"
    return a


def workload_name(a):
    return (f"C3 synthetic train step: per-GPU batch {a.batch}, L={a.seq_len}, d={D}, hid={HID}, C={1 + a.neg}, "
            f"V={V_ITEMS}, SASRec+ItC(ts2=0.4){'+DR' if a.dr else ''}, dropout 0.5, {'uniform ids (roofline variant)' if a.ids == 'uniform' else 'realistic ids (median-5 histories, left-padded)'}, "
            f"{'exact-fp32 path' if a.precision == 'fp32' else 'tcgen05 ' + a.precision.upper() + ' GEMM stages, fp32 accumulate/storage'}")


REALISTIC = False


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


def cpu_reference_steps(a, steps, warmup, sample):
    """Times `steps` CPU train steps on `sample` sequences of the workload; returns seq/s."""
    from common import make_keep_masks, make_params
    from oracle import amid_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    L, C = a.seq_len, 1 + a.neg
    rng = np.random.default_rng(1234)
    P = {k: v.requires_grad_(True) for k, v in make_params(7, V_ITEMS, D, L, HID, sample, isDR=a.dr).items()}
    opt = torch.optim.Adam(list(P.values()), lr=5e-4)              # train_sr.py:480
    times = []
    for it in range(warmup + steps):
        b = synth_batch(rng, sample, L, C, V_ITEMS)
        t0 = time.perf_counter()
        masks = make_keep_masks(it, sample, L, D)                   # F.dropout's Bernoulli draws
        outs = O.sasrec_forward(P, b["i_node"], b["neg_samples"], b["seq_d1"], b["seq_d2"], isInC=False, isItC=True,
                                ts1=0.5, ts2=0.4, isDR=a.dr, masks=masks, closed_form=True)
        loss = O.loss_cls(outs[0], outs[1], b["label"], b["domain_id"])
        if a.dr:
            loss = loss + 0.01 * O.loss_dr_e(*outs, b["label"], b["domain_id"])
        opt.zero_grad()
        loss.backward()
        opt.step()                                                   # dense Adam over the [V,128] table too
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    return sample / (ms / 1e3), ms, cores


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, ms, cores = cpu_reference_steps(a, a.steps, a.warmup, a.cpu_sample)
    sample = (f"{a.cpu_sample} sequences per step of the same workload (L={a.seq_len}); oracle port of the reference "
              f"path + torch.optim.Adam (dense table), closed-form ItC (the literal [bs,bs,n,n] ItC needs 275 GB here)")
    line = {
        "impl": "reference", "metric": "train_seqs_per_sec", "value": value, "unit": "seq/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": workload_name(a)},
        "cpu_baseline": {"value": value, "unit": "seq/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "seq/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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
        "k_attn_fwd": {"flop": attn_full, "byte": 4 * act}, "k_attn_bwd": {"flop": 2.5 * attn_full, "byte": 8 * act},
        "k_attn_fwd_mma": {"flop": attn_full, "byte": 4 * act}, "k_attn_bwd_mma": {"flop": 2.5 * attn_full, "byte": 8 * act},
        "k_attn_fwd_mma3": {"flop": attn_full, "byte": 4 * act}, "k_attn_bwd_mma3": {"flop": 2.5 * attn_full, "byte": 8 * act},
        "k_attn_fwd_tc": {"flop": attn_full, "byte": 4 * act}, "k_attn_bwd_tc": {"flop": 2.5 * attn_full, "byte": 8 * act},
        "k_seq_embed": {"byte": rows_seq * (2 * D * 4 + 8)}, "k_gather": {"byte": rows_items * (2 * D * 4 + 8)},
        "k_embed_all": {"byte": (2 * rows_seq + rows_items) * (2 * D * 4 + 8)},
        "k_mim_scores": {"flop": 2.0 * B * L * L * D, "byte": 2 * act},
        "k_mim_scores_mma": {"flop": 2.0 * B * L * L * D, "byte": 2 * act},
        "k_mim_scores_tc5": {"flop": 3 * 2.0 * B * L * L * D, "byte": 2 * act},   # 3xTF32: three MMAs per product
    }
    for k, (f, b) in chain.items():
        for suffix in ("", "_tc", "_16", "_x3"):
            table[k + suffix] = {"flop": f, "byte": b}
    return table.get(name)


def measured_traffic(name, a):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed `ncu --set full` capture
    (profiles/r01_traffic_c3.json); only meaningful for the default C3 shape it was captured on."""
    if (a.batch, a.seq_len, a.neg) != (1024, 200, 1):
        return None
    path = os.path.join(ROOT, "profiles", "r01_traffic_c3.json")
    if not os.path.exists(path):
        return None
    return json.load(open(path)).get("bytes_per_launch", {}).get(name)


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


def eval_full_catalogue_users_per_sec(a, n_batches=12):
    """BASELINE config 5 on one GPU: 256-user eval batches (L=20, the C1 eval shape), every user ranked against the
    whole item pool of its target domain (16,084 / 12,153 items: the cloth_sport_train75 pools) by
    csrc/catalogue.cu; includes the encoder forward, the D2H of the counts and the metric reduce."""
    from amid_b200 import evaluate
    from amid_b200.engine import Trainer
    from amid_b200.model_seq import SASRec
    B, L = 256, 20
    n1, n2 = 16084, 12153
    torch.manual_seed(2)
    m = SASRec(0, D, V_ITEMS, D, L, HID, B, False, True, 0.5, 0.4, isDR=False).cuda().eval()
    m.cfg.precision = a.precision
    tr = Trainer(m)
    rng = np.random.default_rng(9)
    perm = torch.from_numpy(rng.permutation(V_ITEMS)[:n1 + n2])
    pool1, pool2 = perm[:n1], perm[n1:]
    cat = tr.catalogue(pool1, pool2)
    bs = []
    for _ in range(4):
        b = synth_batch(rng, B, L, 2, V_ITEMS)
        dom = b["domain_id"]
        b["i_node"] = torch.where(dom == 0, pool1[torch.from_numpy(rng.integers(0, n1, B))], pool2[torch.from_numpy(rng.integers(0, n2, B))])
        b["overlap_label"] = torch.from_numpy(rng.integers(0, 2, B))
        bs.append(tr.to_device(b))
    evaluate.evaluate_full_catalogue(tr.P, tr.cfg, cat, bs[:2])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    evaluate.evaluate_full_catalogue(tr.P, tr.cfg, cat, [bs[i % 4] for i in range(n_batches)])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    pairs = n_batches * B * (n1 + n2) / 2.0
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
    return {"metric": "eval_users_per_sec_full_catalogue", "value": n_batches * B / dt, "unit": "users/s",
            "ms_per_batch": 1e3 * dt / n_batches, "pairs_per_sec": pairs / dt,
            "kernel_pairs_per_sec": kernel_pairs, "kernel_users_per_sec": kernel_pairs / n1,
            "config": f"256 users per batch, L={L}, pools {n1} / {n2} items (every pool item scored, fp32 post-sigmoid), 1 GPU"}


def eval_users_per_sec(a, n_batches=20):
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
    bs = [tr.to_device(synth_batch(rng, B, L, C, V_ITEMS)) for _ in range(4)]

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
    dt = time.perf_counter() - t0
    return {"metric": "eval_users_per_sec", "value": n_batches * B / dt, "unit": "users/s", "ms_per_batch": 1e3 * dt / n_batches,
            "config": f"{B} users x {C} candidates per batch, L={L}, SASRec+ItC, device ranking + HR/NDCG/MRR, 1 GPU"}


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
    clk = clocks.stop() if rank == 0 else None
    for i in range(2):
        host_step(i)
    ms_e2e = timed(host_step, a.steps)
    h2d = sum(v.numel() * (4 if k == "label" else 8) for k, v in host[0].items())

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
    # over the time of the kernels that hold those contractions
    gemm_ms = sum(e["ms_per_step"] for e in breakdown
                  if e["kernel"].split("_tc")[0].split("_16")[0].split("_x3")[0] in ("k_ln_qkv", "k_proj_ffn", "k_ffn_bwd", "k_qkv_bwd", "k_wgrad"))
    gemm_flop = 3.0 * 4 * 12 * L * D * D * B          # 2 blocks x 2 encoders, fwd + dX + dW
    tensor_pipe = None if gemm_ms <= 0 else {
        "flop_per_step": gemm_flop, "ms_in_gemm_kernels": gemm_ms, "achieved_tflops": gemm_flop / (gemm_ms / 1e3) / 1e12,
        "peak_tflops": pk["tensor"], "frac": gemm_flop / (gemm_ms / 1e3) / 1e12 / pk["tensor"],
        "note": "d=128 chains move 5-11 fp32 activation tensors per 3-6 GEMMs (32-64 flop/B): HBM-bound, see per-kernel frac"}
    line = {
        "metric": "train_seqs_per_sec", "value": Bg / (ms_step / 1e3), "unit": "seq/s", "n_gpus": world,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"fp32": "f32", "x3": "f32", "tf32": "tf32", "bf16": "bf16"}[a.precision], "data": "synthetic",
        "config": {"workload": workload_name(a), "global_batch": Bg, "seq_len": L, "parallelism": f"dp{world}",
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
        "kernel_breakdown": breakdown[:12],
        "kernel_tail_ms": {e["kernel"]: round(e["ms_per_step"], 4) for e in breakdown[12:]},
        "final_loss": final_loss,
    }
    try:
        line["eval"] = eval_users_per_sec(a)
    except Exception as e:
        line["eval"] = {"value": None, "error": repr(e)}
    try:
        line["eval_full_catalogue"] = eval_full_catalogue_users_per_sec(a)
    except Exception as e:
        line["eval_full_catalogue"] = {"value": None, "error": repr(e)}
    if world == 1 and not a.no_cpu_baseline:
        try:
            v, ms, cores = cpu_reference_steps(a, 2, 1, a.cpu_sample)
            line["cpu_baseline"] = {"value": v, "unit": "seq/s", "cores": cores, "kind": "port", "ms_per_step": ms,
                                    "sample": f"{a.cpu_sample} sequences per step of the same workload, 1 warm-up + 2 "
                                              f"timed steps; oracle port + torch.optim.Adam, closed-form ItC"}
        except Exception as e:  # the baseline is reported beside the number, never a reason to lose it
            line["cpu_baseline"] = {"value": None, "unit": "seq/s", "cores": os.cpu_count(), "kind": "port",
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

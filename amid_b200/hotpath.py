"""Host-side sequencing of the SASRec hot path over the C ABI (include/amid_b200.h).

``forward`` / ``backward`` mirror SASRec.forward of the reference (model_seq.py:416-443)
and its autograd, kernel by kernel.  PyTorch is used for device memory and streams only;
every arithmetic step is a call into libamid_b200.so.  No CPU fallback exists.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import _abi
from ._abi import Dropout, EncoderSaved, EncoderTensors, HeadTensors, call

D = 128
HEADS = 8


@dataclass
class Config:
    """Constructor arguments of model_seq.SASRec that shape the computation."""
    item_length: int
    seq_len: int
    hid_dim: int
    bs: int                      # the GLOBAL batch baked into trans_bs (model_seq.py:480)
    isInC: bool
    isItC: bool
    ts1: float
    ts2: float
    isDR: bool
    drop_p: float = 0.5          # model_seq.py:335,350,355 (hard-coded in the reference)
    overlap_encoders: bool = True   # run the two domain encoders on two streams
    # GEMM stages: "fp32" exact CUDA-core tiles | "x3" tcgen05 split-operand tiles at fp32-level accuracy (same
    # tolerances as "fp32") | "tf32" / "bf16" single-pass tcgen05 tiles (reduced precision, stated tolerances)
    precision: str = "fp32"

    @property
    def enc_len(self) -> int:    # model_seq.py:399-400
        return self.seq_len * 2 if self.isInC else self.seq_len

    @property
    def head_names(self):
        return ["predictModule"] + (["predict_ips", "predict_gfunc"] if self.isDR else [])


_ENC_FWD = {"fp32": "amid_encoder_fwd", "x3": "amid_encoder_fwd_x3", "tf32": "amid_encoder_fwd_tc",
            "bf16": "amid_encoder_fwd_bf16"}
_ENC_BWD = {"fp32": "amid_encoder_bwd", "x3": "amid_encoder_bwd_x3", "tf32": "amid_encoder_bwd_tc",
            "bf16": "amid_encoder_bwd_bf16"}


class DistCtx:
    """Batch-sharded data parallelism: this rank owns batch positions [j0, j0+B_local) of
    the global batch.  ``group`` is a torch.distributed group."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.sharded = None          # amid_b200.sharded.ShardedTable when the item table is row-sharded (config 4)

    def all_gather_into(self, out, inp):
        self.dist.all_gather_into_tensor(out, inp, group=self.group)

    def all_reduce(self, t):
        self.dist.all_reduce(t, group=self.group)


_SIDE_STREAMS = {}


def _side_stream(dev) -> "torch.cuda.Stream":
    """Second stream per device: the two domain encoders are independent (sac1 / sac2, model_seq.py:425-426),
    so their kernel chains run concurrently and fill each other's wave tails."""
    key = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=key)
    return _SIDE_STREAMS[key]


class _Fork:
    """Run branch 1 on the side stream while branch 0 stays on the current stream; join at exit.
    Every buffer is allocated on the current stream BEFORE the fork (caching-allocator safety)."""

    def __init__(self, dev, enabled: bool):
        self.enabled = enabled
        self.main = torch.cuda.current_stream()
        self.side = _side_stream(dev) if enabled else None

    def __enter__(self):
        if self.enabled:
            ev = torch.cuda.Event()
            ev.record(self.main)
            self.side.wait_event(ev)
        return self

    def branch(self, k: int):
        import contextlib
        if self.enabled and k == 1:
            return torch.cuda.stream(self.side)
        return contextlib.nullcontext()

    def __exit__(self, *exc):
        if self.enabled:
            ev = torch.cuda.Event()
            ev.record(self.side)
            self.main.wait_event(ev)
        return False


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise _abi.AmidError(f"{name} must be a CUDA tensor (there is no CPU path)")
    if t.dtype != dtype:
        raise _abi.AmidError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _abi.AmidError(f"{name} must be contiguous")
    return t


def encoder_struct(P: Dict[str, torch.Tensor], prefix: str) -> EncoderTensors:
    """Fill the C struct from reference-named tensors."""
    s = EncoderTensors()
    s.pos_emb = P[prefix + "pos_emb.weight"].data_ptr()
    for i in range(2):
        s.ln1_w[i] = P[f"{prefix}attention_layernorms.{i}.weight"].data_ptr()
        s.ln1_b[i] = P[f"{prefix}attention_layernorms.{i}.bias"].data_ptr()
        s.in_w[i] = P[f"{prefix}attention_layers.{i}.in_proj_weight"].data_ptr()
        s.in_b[i] = P[f"{prefix}attention_layers.{i}.in_proj_bias"].data_ptr()
        s.out_w[i] = P[f"{prefix}attention_layers.{i}.out_proj.weight"].data_ptr()
        s.out_b[i] = P[f"{prefix}attention_layers.{i}.out_proj.bias"].data_ptr()
        s.ln2_w[i] = P[f"{prefix}forward_layernorms.{i}.weight"].data_ptr()
        s.ln2_b[i] = P[f"{prefix}forward_layernorms.{i}.bias"].data_ptr()
        s.c1_w[i] = P[f"{prefix}forward_layers.{i}.conv1.weight"].data_ptr()
        s.c1_b[i] = P[f"{prefix}forward_layers.{i}.conv1.bias"].data_ptr()
        s.c2_w[i] = P[f"{prefix}forward_layers.{i}.conv2.weight"].data_ptr()
        s.c2_b[i] = P[f"{prefix}forward_layers.{i}.conv2.bias"].data_ptr()
    s.ln3_w = P[prefix + "last_layernorm.weight"].data_ptr()
    s.ln3_b = P[prefix + "last_layernorm.bias"].data_ptr()
    return s


def heads_struct(P: Dict[str, torch.Tensor], names):
    arr = (HeadTensors * len(names))()
    for i, h in enumerate(names):
        arr[i].w0 = P[h + ".fc.0.weight"].data_ptr()
        arr[i].b0 = P[h + ".fc.0.bias"].data_ptr()
        arr[i].w2 = P[h + ".fc.2.weight"].data_ptr()
        arr[i].b2 = P[h + ".fc.2.bias"].data_ptr()
    return arr


class _Saved:
    """Activation buffers of one encoder pass (amid_encoder_saved)."""

    def __init__(self, B: int, L: int, dev):
        M = B * L
        f = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        self.big = f(2, 9, M, D)                  # qn q k v o x1 y h xout per block
        self.lse = f(2, B * HEADS * L)
        self.st = f(5, M, 2)
        s = EncoderSaved()
        for i in range(2):
            for j, n in enumerate(("qn", "q", "k", "v", "o", "x1", "y", "h", "xout")):
                getattr(s, n)[i] = self.big[i, j].data_ptr()
            s.lse[i] = self.lse[i].data_ptr()
            s.st1[i] = self.st[i].data_ptr()
            s.st2[i] = self.st[2 + i].data_ptr()
        s.st3 = self.st[4].data_ptr()
        self.struct = s


class _Mim:
    """State of one InterComp/InnerComp application (closed form)."""
    __slots__ = ("name", "p", "gate", "coef", "active", "n_active", "scal", "Ssum", "E", "esum", "other", "n")


def _dropout(cfg: Config, train: bool, seed: int, site_base: int, batch_offset: int = 0) -> Dropout:
    """batch_offset = position of this rank's first sample in the global batch: the keep bits are indexed by global
    sample, so data-parallel training with dropout equals single-GPU training on the global batch."""
    return Dropout(1 if (train and cfg.drop_p > 0) else 0, float(cfg.drop_p), int(seed) & (2**64 - 1), site_base,
                   int(batch_offset))


def _mim_forward(P, name: str, m_global: torch.Tensor, other: torch.Tensor, n: int, ts: float, j0: int,
                 dist: Optional[DistCtx], want_esum: bool) -> _Mim:
    """gate + aggregate + project for one direction; `other` is the local [B,n,128] tensor
    that gets aggregated (seq_d2 features for itc_d1, model_seq.py:483-497)."""
    dev = other.device
    Bg = m_global.numel()
    Bl = other.numel() // (n * D)
    w_bs = P[name + ".trans_bs.weight"]
    if w_bs.numel() != Bg:
        raise _abi.AmidError(f"{name}: batch {Bg} != bs {w_bs.numel()} baked into trans_bs (model_seq.py:480); "
                             "the reference requires drop_last batches of exactly bs rows")
    st = _Mim()
    st.name, st.other, st.n = name, other, n
    f = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
    st.p, st.gate, st.coef = f(Bg), f(Bg), f(Bg)
    st.active = torch.empty(Bg, device=dev, dtype=torch.int32)
    st.n_active = torch.empty(1, device=dev, dtype=torch.int32)
    st.scal = f(1)
    st.Ssum, st.E = f(n, D), f(n, D)
    st.esum = f(D) if want_esum else None
    s = _stream()
    call("amid_mim_gate", _ptr(m_global), _ptr(w_bs), Bg, float(ts), _ptr(st.p), _ptr(st.gate), _ptr(st.coef),
         _ptr(st.active), _ptr(st.n_active), _ptr(st.scal), s)
    call("amid_mim_aggregate", _ptr(other), _ptr(st.coef), _ptr(st.active), _ptr(st.n_active), j0, Bl, n,
         _ptr(st.Ssum), s)
    if dist is not None and dist.world > 1:
        dist.all_reduce(st.Ssum)
    call("amid_mim_project", _ptr(st.Ssum), _ptr(P[name + ".trans_nn.weight"]), _ptr(P[name + ".trans_nn.bias"]),
         _ptr(P[name + ".trans_bs.bias"]), _ptr(st.scal), n, _ptr(st.E), _ptr(st.esum), s)
    return st


def _mim_scores(a: torch.Tensor, b: torch.Tensor, n: int, dist: Optional[DistCtx], tc: bool = False) -> torch.Tensor:
    Bl = a.numel() // (n * D)
    m = torch.empty(Bl, device=a.device, dtype=torch.float32)
    call("amid_mim_scores_tc" if tc else "amid_mim_scores", _ptr(a), _ptr(b), Bl, n, _ptr(m), _stream())
    if dist is not None and dist.world > 1:
        mg = torch.empty(Bl * dist.world, device=a.device, dtype=torch.float32)
        dist.all_gather_into(mg, m)
        return mg
    return m


def _mim_backward(P, G, st: _Mim, dE: torch.Tensor, d_other: torch.Tensor, j0: int, dist: Optional[DistCtx]):
    dev = dE.device
    Bl = st.other.numel() // (st.n * D)
    ws = torch.empty(st.n, D, device=dev, dtype=torch.float32)
    dw_bs_local = torch.empty(Bl, device=dev, dtype=torch.float32)
    call("amid_mim_bwd", _ptr(dE), _ptr(st.Ssum), _ptr(st.other), _ptr(P[st.name + ".trans_nn.weight"]),
         _ptr(P[st.name + ".trans_nn.bias"]), _ptr(st.coef), _ptr(st.gate), _ptr(st.active), _ptr(st.n_active),
         _ptr(st.scal), j0, Bl, st.n, _ptr(G[st.name + ".trans_nn.weight"]), _ptr(G[st.name + ".trans_nn.bias"]),
         _ptr(G[st.name + ".trans_bs.bias"]), _ptr(dw_bs_local), _ptr(d_other), _ptr(ws), _stream())
    gw = G[st.name + ".trans_bs.weight"]
    if dist is not None and dist.world > 1:
        # dE is already the global sum, so trans_nn / bias grads are complete on every rank;
        # pre-divide so the later dense all-reduce (sum) leaves them unchanged.  The w_bs
        # slices are disjoint per rank.
        for k in (".trans_nn.weight", ".trans_nn.bias", ".trans_bs.bias"):
            G[st.name + k].mul_(1.0 / dist.world)
        gw.zero_()
        gw.view(-1)[j0:j0 + Bl].copy_(dw_bs_local)
    else:
        gw.view(-1).copy_(dw_bs_local)


class Ctx:
    """Everything backward needs from a forward pass."""
    pass


def forward(P: Dict[str, torch.Tensor], cfg: Config, i_node, neg_samples, seq_d1, seq_d2, *, train: bool,
            seed: int = 0, dist: Optional[DistCtx] = None, need_ctx: bool = True):
    """SASRec.forward (model_seq.py:416-443).  Returns (probs [n_heads,2,B,C], ctx)."""
    table = _chk(P["item_emb_layer.emb_item.weight"], torch.float32, "item table")
    V = table.shape[0]
    if table.shape[1] != D:
        raise _abi.AmidError(f"item_emb_dim must be {D} (got {table.shape[1]})")
    dev = table.device
    B, L = seq_d1.shape
    if L != cfg.seq_len:
        raise _abi.AmidError(f"sequence length {L} != seq_len {cfg.seq_len} of the positional table")
    ids_items = torch.cat((i_node.reshape(B, 1), neg_samples.reshape(B, -1)), 1).contiguous()
    Cn = ids_items.shape[1]
    seqs = [_chk(seq_d1, torch.int64, "seq_d1"), _chk(seq_d2, torch.int64, "seq_d2")]
    _chk(ids_items, torch.int64, "i_node/neg_samples")
    Le = cfg.enc_len
    world = dist.world if dist is not None else 1
    j0 = dist.rank * B if dist is not None else 0
    s = _stream()
    f = lambda *sh: torch.empty(*sh, device=dev, dtype=torch.float32)

    ctx = Ctx()
    ctx.route = None
    sharded = dist.sharded if dist is not None else None
    if sharded is not None:
        # row-sharded table: one all-to-all lookup fetches the unique rows of the step into a compact step table;
        # from here on the kernels see (step table, position -> step-table row) instead of (table, item id)
        route = sharded.lookup(torch.cat((ids_items.reshape(-1), seqs[0].reshape(-1), seqs[1].reshape(-1))))
        table, V = route.rows, route.rows.shape[0]
        n_it, n_sq = B * Cn, B * L
        ids_items = route.virtual_ids[:n_it].view(B, Cn)
        seqs = [route.virtual_ids[n_it:n_it + n_sq].view(B, L), route.virtual_ids[n_it + n_sq:].view(B, L)]
        ctx.route = route
    ctx.B, ctx.L, ctx.Le, ctx.C, ctx.V, ctx.train, ctx.seed, ctx.j0 = B, L, Le, Cn, V, train, seed, j0
    ctx.ids_items, ctx.seqs = ids_items, seqs

    # a1 (+a2): every table read of the step
    items = f(B, Cn, D)
    ctx.items = items
    encs, x0s, tms, saveds, incs, raws = [], [], [], [], [], []
    pre_x0 = pre_tm = None
    if not cfg.isInC:
        pre_x0 = [f(B * Le, D), f(B * Le, D)]
        pre_tm = [torch.empty(B * Le * 4, device=dev, dtype=torch.int32) for _ in range(2)]
        drop0 = _dropout(cfg, train, seed, 0, j0)
        call("amid_embed_all_fwd", _ptr(table), V, _ptr(ids_items), B * Cn, _ptr(seqs[0]), _ptr(seqs[1]),
             _ptr(P["sac1.pos_emb.weight"]), _ptr(P["sac2.pos_emb.weight"]), B, L, _ptr(items), _ptr(pre_x0[0]),
             _ptr(pre_x0[1]), _ptr(pre_tm[0]), _ptr(pre_tm[1]), C.byref(drop0), s)
    else:
        call("amid_emb_gather_fwd", _ptr(table), V, _ptr(ids_items), B * Cn, _ptr(items), s)
    for k, sac in enumerate(("sac1.", "sac2.")):
        drop = _dropout(cfg, train, seed, 8 * k, j0)
        if cfg.isInC:
            x0 = f(B * Le, D)
            tm = torch.empty(B * Le * 4, device=dev, dtype=torch.int32)
            # model_seq.py:422-424  InnerComp before the encoder
            raw = f(B, L, D)
            call("amid_emb_gather_fwd", _ptr(table), V, _ptr(seqs[k]), B * L, _ptr(raw), s)
            mg = _mim_scores(raw, raw, L, dist, cfg.precision != "fp32")
            st = _mim_forward(P, f"inc_d{k + 1}", mg, raw, L, cfg.ts1, j0, dist, want_esum=False)
            cat = f(B, 2 * L, D)
            call("amid_mim_concat", _ptr(raw), _ptr(st.E), B, L, _ptr(cat), s)
            call("amid_seq_embed_fwd", None, V, None, _ptr(cat), _ptr(P[sac + "pos_emb.weight"]), B, Le, _ptr(x0),
                 _ptr(tm), C.byref(drop), s)
            incs.append(st)
            raws.append(raw)
        else:
            x0, tm = pre_x0[k], pre_tm[k]
        x0s.append(x0); tms.append(tm)
    ws_bytes = _abi.lib().amid_encoder_fwd_workspace_bytes(B, Le)
    wss = [torch.empty(ws_bytes, device=dev, dtype=torch.uint8) for _ in range(2)]
    for k in range(2):
        saveds.append(_Saved(B, Le, dev))
        encs.append(f(B * Le, D))
    with _Fork(dev, cfg.overlap_encoders) as fork:
        for k, sac in enumerate(("sac1.", "sac2.")):
            with fork.branch(k):
                drop = _dropout(cfg, train, seed, 8 * k, j0)
                es = encoder_struct(P, sac)
                call(_ENC_FWD[cfg.precision], C.byref(es), _ptr(x0s[k]), _ptr(tms[k]), B, Le, C.byref(drop),
                     C.byref(saveds[k].struct), _ptr(encs[k]), _ptr(wss[k]), ws_bytes, _stream())
    ctx.encs, ctx.x0s, ctx.tms, ctx.saveds, ctx.incs, ctx.raws = encs, x0s, tms, saveds, incs, raws

    # a6 + a7: ItC and mean pool
    us = [f(B, D), f(B, D)]
    ctx.itcs = []
    if cfg.isItC:
        mg = _mim_scores(encs[0], encs[1], Le, dist, cfg.precision != "fp32")   # max of a matrix == max of its transpose
        for k in range(2):
            st = _mim_forward(P, f"itc_d{k + 1}", mg, encs[1 - k].view(B, Le, D), Le, cfg.ts2, j0, dist, True)
            ctx.itcs.append(st)
            call("amid_meanpool_fwd", _ptr(encs[k]), _ptr(st.esum), B, Le, float(2 * Le), _ptr(us[k]), s)
    else:
        for k in range(2):
            call("amid_meanpool_fwd", _ptr(encs[k]), None, B, Le, float(Le), _ptr(us[k]), s)
    ctx.us = us

    # a8: scorer heads
    names = cfg.head_names
    nh = len(names)
    probs = f(nh, 2, B, Cn)
    hs = heads_struct(P, names)
    call("amid_score_fwd", _ptr(us[0]), _ptr(us[1]), _ptr(items), hs, nh, cfg.hid_dim, B, Cn, _ptr(probs), s)
    ctx.probs = probs
    ctx.world = world
    return probs, (ctx if need_ctx else None)


def grad_buffers(P: Dict[str, torch.Tensor], skip=("item_emb_layer.emb_item.weight",)) -> Dict[str, torch.Tensor]:
    """One flat fp32 buffer with a view per dense parameter (zero-initialised)."""
    names = [n for n in P if n not in skip]
    total = sum((P[n].numel() + 3) // 4 * 4 for n in names)
    flat = torch.zeros(total, device=next(iter(P.values())).device, dtype=torch.float32)
    out, off = {}, 0
    for n in names:
        k = P[n].numel()
        out[n] = flat[off:off + k].view(P[n].shape)
        off += (k + 3) // 4 * 4
    out["__flat__"] = flat
    return out


def backward(P: Dict[str, torch.Tensor], cfg: Config, ctx: Ctx, dprobs: torch.Tensor,
             G: Optional[Dict[str, torch.Tensor]] = None, dist: Optional[DistCtx] = None):
    """Backward of ``forward``.  Returns (G, ids_all, grad_rows_all): dense parameter
    gradients by reference name, and the per-row table gradients (not yet reduced)."""
    B, L, Le, Cn = ctx.B, ctx.L, ctx.Le, ctx.C
    dev = dprobs.device
    s = _stream()
    f = lambda *sh: torch.empty(*sh, device=dev, dtype=torch.float32)
    if G is None:
        G = grad_buffers(P)
    names = cfg.head_names
    nh = len(names)
    _chk(dprobs, torch.float32, "dprobs")

    # one buffer for every table-gradient row: [items | seq_d1 | seq_d2]
    n_items, n_seq = B * Cn, B * L
    rows_all = f(n_items + 2 * n_seq, D)
    ditems = rows_all[:n_items]
    dseq = [rows_all[n_items:n_items + n_seq], rows_all[n_items + n_seq:]]
    ids_all = torch.cat((ctx.ids_items.reshape(-1), ctx.seqs[0].reshape(-1), ctx.seqs[1].reshape(-1)))

    # a8 backward
    dus = [f(B, D), f(B, D)]
    hs = heads_struct(P, names)
    gs = heads_struct(G, names)
    wsb = _abi.lib().amid_score_bwd_workspace_bytes(nh, cfg.hid_dim, B, Cn)
    ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
    call("amid_score_bwd", _ptr(ctx.us[0]), _ptr(ctx.us[1]), _ptr(ctx.items), hs, nh, cfg.hid_dim, B, Cn,
         _ptr(ctx.probs), _ptr(dprobs), _ptr(dus[0]), _ptr(dus[1]), _ptr(ditems), gs, _ptr(ws), wsb, s)

    # a7 / a6 backward
    d_encs = [f(B * Le, D), f(B * Le, D)]
    if cfg.isItC:
        dcols = [f(D), f(D)]
        for k in range(2):
            call("amid_meanpool_bwd", _ptr(dus[k]), B, Le, float(2 * Le), 0, _ptr(d_encs[k]), _ptr(dcols[k]), s)
        for k in range(2):
            if dist is not None and dist.world > 1:
                dist.all_reduce(dcols[k])
            dE = dcols[k].view(1, D).expand(Le, D).contiguous()   # every row of dE is the same vector
            _mim_backward(P, G, ctx.itcs[k], dE, d_encs[1 - k], ctx.j0, dist)
    else:
        for k in range(2):
            call("amid_meanpool_bwd", _ptr(dus[k]), B, Le, float(Le), 0, _ptr(d_encs[k]), None, s)

    # a3-a5 backward, then a2 / a1
    wsb = _abi.lib().amid_encoder_bwd_workspace_bytes(B, Le)
    wss = [torch.empty(wsb, device=dev, dtype=torch.uint8) for _ in range(2)]
    dx0s = [f(B * Le, D) if cfg.isInC else dseq[k] for k in range(2)]
    with _Fork(dev, cfg.overlap_encoders and not cfg.isInC) as fork:
        for k, sac in enumerate(("sac1.", "sac2.")):
            with fork.branch(k):
                drop = _dropout(cfg, ctx.train, ctx.seed, 8 * k, ctx.j0)
                es = encoder_struct(P, sac)
                gstruct = encoder_struct(G, sac)
                dx0 = dx0s[k]
                call(_ENC_BWD[cfg.precision], C.byref(es), _ptr(ctx.x0s[k]), _ptr(ctx.tms[k]), B, Le, C.byref(drop),
                     C.byref(ctx.saveds[k].struct), _ptr(ctx.encs[k]), _ptr(d_encs[k]), C.byref(gstruct), _ptr(dx0),
                     _ptr(wss[k]), wsb, _stream())
                gpos = G[sac + "pos_emb.weight"]
                call("amid_seq_embed_bwd", _ptr(dx0), _ptr(ctx.tms[k]), B, Le, _ptr(gpos), C.byref(drop), _stream())
                if cfg.isInC:
                    # dx0 is the gradient of cat(seq, E): the E half summed over the batch is exactly
                    # what amid_seq_embed_bwd just accumulated into rows [L, 2L) of the positional grad.
                    st = ctx.incs[k]
                    dE = gpos[L:2 * L].clone()
                    if dist is not None and dist.world > 1:
                        dist.all_reduce(dE)
                    dseq[k].view(B, L, D).copy_(dx0.view(B, 2 * L, D)[:, :L])
                    _mim_backward(P, G, st, dE, dseq[k], ctx.j0, dist)
    return G, ids_all, rows_all


def segreduce(ids_all: torch.Tensor, rows_all: torch.Tensor, V: int):
    """Deterministic sort-by-index segmented reduction of the table-gradient rows."""
    n = ids_all.numel()
    dev = rows_all.device
    uniq_ids = torch.empty(n, device=dev, dtype=torch.int64)
    uniq_grads = torch.empty(n, D, device=dev, dtype=torch.float32)
    n_uniq = torch.zeros(1, device=dev, dtype=torch.int32)
    wsb = _abi.lib().amid_embgrad_workspace_bytes(n)
    ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
    call("amid_embgrad_segreduce", _ptr(ids_all), _ptr(rows_all), n, V, _ptr(uniq_ids), _ptr(uniq_grads), _ptr(n_uniq),
         _ptr(ws), wsb, _stream())
    return uniq_ids, uniq_grads, n_uniq


def dense_table_grad(uniq_ids, uniq_grads, n_uniq, V: int) -> torch.Tensor:
    """The dense [V,128] gradient torch.optim.Adam expects from the drop-in module."""
    dense = torch.zeros(V, D, device=uniq_grads.device, dtype=torch.float32)
    call("amid_embgrad_scatter_dense", _ptr(uniq_ids), _ptr(uniq_grads), _ptr(n_uniq), uniq_ids.numel(), _ptr(dense), V,
         _stream())
    return dense


def loss_fwd_bwd(probs: torch.Tensor, labels: torch.Tensor, domain_id: torch.Tensor, ob_label, mode: int,
                 dr_e_w: float, global_batch: int):
    """Fused losses + gradient w.r.t. the probabilities (amid_loss_fwd_bwd)."""
    nh, _, B, Cn = probs.shape
    losses = torch.empty(3, device=probs.device, dtype=torch.float32)
    dprobs = torch.empty_like(probs)
    call("amid_loss_fwd_bwd", _ptr(probs), nh, B, Cn, _ptr(_chk(labels, torch.float32, "labels")),
         _ptr(_chk(domain_id, torch.int64, "domain_id")), _ptr(ob_label), mode, float(dr_e_w),
         1.0 / float(global_batch * Cn), _ptr(losses), _ptr(dprobs), _stream())
    return losses, dprobs


def dropout_masks(cfg: Config, B: int, Le: int, seed: int, dev) -> dict:
    """The keep-masks the kernels use for (seed) -- test support for oracle mask injection."""
    out = {}
    for k, sac in enumerate(("sac1", "sac2")):
        drop = _dropout(cfg, True, seed, 8 * k)
        m = {}

        def feat(site):
            t = torch.empty(B * Le * D, device=dev, dtype=torch.uint8)
            call("amid_dropout_mask_feature", C.byref(drop), 8 * k + site, B * Le, _ptr(t), _stream())
            return t.view(B, Le, D).bool()

        def attn(site):
            t = torch.empty(B * HEADS * Le * Le, device=dev, dtype=torch.uint8)
            call("amid_dropout_mask_attn", C.byref(drop), 8 * k + site, B, Le, _ptr(t), _stream())
            return t.view(B, HEADS, Le, Le).bool()

        m["emb"] = feat(0)
        for i in range(2):
            m[f"attn{i}"] = attn(1 + 3 * i)
            m[f"ffn1_{i}"] = feat(2 + 3 * i)
            m[f"ffn2_{i}"] = feat(3 + 3 * i)
        out[sac] = m
    return out

"""Drop-in replacement for the reference's ``model_seq.SASRec`` (model_seq.py:390-443).

Same constructor signature, same ``forward`` signature and return values, same
``state_dict`` names, ordinary leaf ``nn.Parameter`` s -- so ``train_sr.py`` /
``train_sr_dr.py`` drive it unchanged (``from model_seq import *``; put
``amid_b200/dropin`` on ``sys.path``).  Underneath, every operation
runs in the hand-written sm_100a kernels of libamid_b200.so through the C ABI; the
module is CUDA-only and raises if the library or a GPU is missing.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import hotpath
from ._abi import AmidError

__all__ = ["SASRec", "embItemLayerEnhance", "predictModule", "PointWiseFeedForward", "Log2feats", "InterComp",
           "InnerComp"]


# ---- parameter containers with the reference's names, shapes and default initialisers.
# Their forward() is never used: SASRec.forward runs the fused hot path.
def _kaiming_uniform_(w: torch.Tensor, fan_in: int):
    bound = 1.0 / math.sqrt(fan_in)      # nn.Linear / nn.Conv1d default: kaiming_uniform(a=sqrt(5))
    with torch.no_grad():
        return w.uniform_(-bound, bound)


class _Linear(nn.Module):
    def __init__(self, fin, fout, conv=False):
        super().__init__()
        self.weight = nn.Parameter(_kaiming_uniform_(torch.empty((fout, fin, 1) if conv else (fout, fin)), fin))
        self.bias = nn.Parameter(_kaiming_uniform_(torch.empty(fout), fin))


class _LayerNorm(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d))
        self.bias = nn.Parameter(torch.zeros(d))


class _MHA(nn.Module):
    """Parameter layout of torch.nn.MultiheadAttention(d, 8, 0.5) (model_seq.py:348-350)."""

    def __init__(self, d):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d, d))
        nn.init.xavier_uniform_(self.in_proj_weight)
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d))
        self.out_proj = _Linear(d, d)
        with torch.no_grad():
            self.out_proj.bias.zero_()


class embItemLayerEnhance(nn.Module):            # model_seq.py:22-29
    def __init__(self, item_length, emb_dim):
        super().__init__()
        self.emb_item = nn.Embedding(item_length, emb_dim)


class predictModule(nn.Module):                  # model_seq.py:32-54
    def __init__(self, emb_dim, hid_dim):
        super().__init__()
        self.fc = nn.ModuleDict({"0": _Linear(emb_dim * 2, hid_dim), "2": _Linear(hid_dim, 1)})


class PointWiseFeedForward(nn.Module):           # model_seq.py:311-326
    def __init__(self, hidden_units, dropout_rate=0.5):
        super().__init__()
        self.conv1 = _Linear(hidden_units, hidden_units, conv=True)
        self.conv2 = _Linear(hidden_units, hidden_units, conv=True)


class Log2feats(nn.Module):                      # model_seq.py:331-357
    def __init__(self, user_length, user_emb_dim, item_length, item_emb_dim, seq_len, hid_dim):
        super().__init__()
        self.pos_emb = nn.Embedding(seq_len, item_emb_dim)
        self.attention_layernorms = nn.ModuleList(_LayerNorm(user_emb_dim) for _ in range(2))
        self.attention_layers = nn.ModuleList(_MHA(user_emb_dim) for _ in range(2))
        self.forward_layernorms = nn.ModuleList(_LayerNorm(user_emb_dim) for _ in range(2))
        self.forward_layers = nn.ModuleList(PointWiseFeedForward(user_emb_dim) for _ in range(2))
        self.last_layernorm = _LayerNorm(user_emb_dim)


class InterComp(nn.Module):                      # model_seq.py:474-481
    def __init__(self, user_emb_dim, bs, threshold):
        super().__init__()
        self.bs, self.threshold = bs, threshold
        self.trans_nn = _Linear(user_emb_dim, user_emb_dim)
        self.trans_bs = _Linear(bs, 1)


class InnerComp(InterComp):                      # model_seq.py:450-457
    pass


class _HotPathFn(torch.autograd.Function):
    """Whole-model autograd node: forward and backward are kernel sequences over the C ABI."""

    @staticmethod
    def forward(ctx, module, need_ctx, i_node, neg_samples, seq_d1, seq_d2, *params):
        P = dict(zip(module._pnames, (p.detach() for p in params)))
        train = module.training
        seed = module._next_seed() if train else 0
        probs, hctx = hotpath.forward(P, module.cfg, i_node, neg_samples, seq_d1, seq_d2, train=train, seed=seed,
                                      dist=module._dist, need_ctx=need_ctx)
        ctx.module, ctx.hctx, ctx.P = module, hctx, P
        nh = probs.shape[0]
        outs = tuple(probs[h, k] for h in range(nh) for k in range(2))
        return outs

    @staticmethod
    def backward(ctx, *douts):
        module, hctx, P = ctx.module, ctx.hctx, ctx.P
        if hctx is None:
            raise AmidError("backward called on a forward that ran under no_grad")
        probs = hctx.probs
        dprobs = torch.stack([torch.zeros_like(probs[0, 0]) if g is None else g.to(torch.float32)
                              for g in douts]).view_as(probs).contiguous()
        G, ids_all, rows_all = hotpath.backward(P, module.cfg, hctx, dprobs, dist=module._dist)
        uid, ug, nu = hotpath.segreduce(ids_all, rows_all, hctx.V)
        module._last_table_grad = (uid, ug, nu)
        grads = []
        for n in module._pnames:
            if n == "item_emb_layer.emb_item.weight":
                grads.append(hotpath.dense_table_grad(uid, ug, nu, hctx.V))
            else:
                grads.append(G[n])
        return (None, None, None, None, None, None, *grads)


class SASRec(nn.Module):
    """model_seq.SASRec with the reference constructor (model_seq.py:391) and forward (:416)."""

    def __init__(self, user_length, user_emb_dim, item_length, item_emb_dim, seq_len, hid_dim, bs, isInC, isItC,
                 threshold1, threshold2, isDR=False):
        super().__init__()
        if item_emb_dim != hotpath.D or user_emb_dim != hotpath.D:
            raise AmidError(f"amid_b200 kernels are specialised for emb_dim = {hotpath.D} (run.sh default)")
        if hid_dim > 64:
            raise AmidError("hid_dim > 64 is not supported by the scorer kernels")
        self.user_emb_dim = user_emb_dim
        self.item_emb_layer = embItemLayerEnhance(item_length, item_emb_dim)
        self.isInC, self.isItC, self.isDR = bool(isInC), bool(isItC), bool(isDR)
        self.cfg = hotpath.Config(item_length=item_length, seq_len=seq_len, hid_dim=hid_dim, bs=bs, isInC=self.isInC,
                                  isItC=self.isItC, ts1=float(threshold1), ts2=float(threshold2), isDR=self.isDR)
        enc_len = seq_len
        if self.isInC:                                         # model_seq.py:398-402
            enc_len = seq_len * 2
            self.inc_d1 = InnerComp(user_emb_dim, bs, threshold1)
            self.inc_d2 = InnerComp(user_emb_dim, bs, threshold1)
        if self.isItC:                                         # model_seq.py:403-405
            self.itc_d1 = InterComp(user_emb_dim, bs, threshold2)
            self.itc_d2 = InterComp(user_emb_dim, bs, threshold2)
        self.sac1 = Log2feats(user_length, user_emb_dim, item_length, item_emb_dim, enc_len, hid_dim)
        self.sac2 = Log2feats(user_length, user_emb_dim, item_length, item_emb_dim, enc_len, hid_dim)
        self.predictModule = predictModule(user_emb_dim, hid_dim)
        if self.isDR:                                          # model_seq.py:410-414
            self.predict_ips = predictModule(user_emb_dim, hid_dim)
            self.predict_gfunc = predictModule(user_emb_dim, hid_dim)
        self._pnames = [n for n, _ in self.named_parameters()]
        self._seed_base = int(torch.initial_seed()) & 0x7FFFFFFFFFFF
        self._step = 0
        self._dist = None
        self._last_table_grad = None

    # per-forward dropout seed (counter-based RNG in the kernels)
    def _next_seed(self) -> int:
        self._step += 1
        return (self._seed_base * 1000003 + self._step) & (2**63 - 1)

    def set_distributed(self, dist_ctx):
        """Batch-sharded data parallelism: see amid_b200.hotpath.DistCtx."""
        self._dist = dist_ctx

    def param_dict(self):
        return {n: p for n, p in self.named_parameters()}

    def forward(self, u_node, i_node, neg_samples, seq_d1, seq_d2, long_tail_mask_d1, long_tail_mask_d2, isTrain=True):
        # u_node, long_tail_mask_*, isTrain are accepted and unused, as in the reference.
        if not seq_d1.is_cuda:
            raise AmidError("amid_b200.SASRec is CUDA-only: move the batch and the module to a B200 (no CPU fallback)")
        params = [p for _, p in self.named_parameters()]
        outs = _HotPathFn.apply(self, torch.is_grad_enabled(), i_node.long().contiguous(), neg_samples.long().contiguous(),
                                seq_d1.long().contiguous(), seq_d2.long().contiguous(), *params)
        return tuple(o.squeeze() for o in outs)               # model_seq.py:54 `.squeeze()`

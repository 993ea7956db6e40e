"""Launch the attention kernels of one implementation a few times (for ncu / timing): python tools/prof_attn.py IMPL B L"""
import ctypes as C
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from amid_b200 import hotpath as hp
from amid_b200._abi import Dropout, call

impl, B, L = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
D, H = 128, 8
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v, dO = (torch.randn(B * L, D, device="cuda", generator=g) for _ in range(4))
o = torch.empty(B * L, D, device="cuda")
lse = torch.empty(B * H * L, device="cuda")
dq, dk, dv = (torch.empty(B * L, D, device="cuda") for _ in range(3))
drop = Dropout(1, 0.5, 12345, 0)


def fwd():
    call("amid_attn_fwd_test", hp._ptr(q), hp._ptr(k), hp._ptr(v), hp._ptr(o), hp._ptr(lse), B, L, C.byref(drop), 1, impl, hp._stream())


def bwd():
    call("amid_attn_bwd_test", hp._ptr(q), hp._ptr(k), hp._ptr(v), hp._ptr(o), hp._ptr(lse), hp._ptr(dO), hp._ptr(dq), hp._ptr(dk),
         hp._ptr(dv), B, L, C.byref(drop), 1, impl, hp._stream())


for fn, name in ((fwd, "fwd"), (bwd, "bwd")):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"impl {impl} {name}: {e0.elapsed_time(e1) / iters * 1e3:.1f} us  (B={B}, L={L})")

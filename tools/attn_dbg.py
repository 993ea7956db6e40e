"""Debug build only (-DAMID_ATTN_DBG): print the clock64 stamps CTA 0 of k_attn_bwd_p recorded."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from amid_b200 import hotpath as hp
from amid_b200._abi import Dropout, call, lib
B, L = 256, 200
D, H = 128, 8
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v, dO = (torch.randn(B * L, D, device="cuda", generator=g) for _ in range(4))
o = torch.empty(B * L, D, device="cuda"); lse = torch.zeros(B * H * L, device="cuda")
dq, dk, dv = (torch.empty(B * L, D, device="cuda") for _ in range(3))
drop = Dropout(1, 0.5, 12345, 0)
for _ in range(2):
    call("amid_attn_bwd_test", hp._ptr(q), hp._ptr(k), hp._ptr(v), hp._ptr(o), hp._ptr(lse), hp._ptr(dO), hp._ptr(dq), hp._ptr(dk),
         hp._ptr(dv), B, L, C.byref(drop), 1, int(sys.argv[1]) if len(sys.argv) > 1 else 4, hp._stream())
torch.cuda.synchronize()
buf = (C.c_longlong * (20 * 256))()
lib().amid_attn_dbg_read(buf)
for w in (10,):
    row = buf[w * 256:(w + 1) * 256]
    t0 = row[1]
    print("warp", w, " ".join(f"{row[i]}:{row[i + 1] - t0}" for i in range(0, 250, 2) if row[i]))

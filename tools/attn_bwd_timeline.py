#!/usr/bin/env python
"""Timeline of the dQ kernel from a -DBW_TIMELINE build (MOLLY_LIB=...): median cycles between clock64 stamps."""
import ctypes, os, statistics, sys
import torch
sys.path.insert(0, ".")
from molly_b200 import _lib
_lib.LIB_PATH = os.environ["MOLLY_LIB"]
from molly_b200 import ops

heads, d, k, n_seq = 20, 64, 1024, 8
h = heads * d
torch.manual_seed(0)
qkv = (torch.randn(n_seq * k, 3 * h, device="cuda") * 0.5).to(torch.bfloat16)
kv_info = torch.tensor([[k, k]] * n_seq, dtype=torch.int32, device="cuda")
mask = torch.ones(n_seq * k, dtype=torch.uint8, device="cuda")
out, lse2 = ops.attention_lse(qkv, n_seq, k, heads, kv_info, mask)
d_out = (torch.randn(n_seq * k, h, device="cuda") * 0.1).to(torch.bfloat16)
for _ in range(3):
    ops.attention_bwd(qkv, out, d_out, lse2, n_seq, k, heads, kv_info, mask)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (16 * 2 * 32 * 8))()
lib = _lib.load()
lib.molly_debug_bw_timeline.argtypes = [ctypes.c_void_p]
assert lib.molly_debug_bw_timeline(buf) == 0
import numpy as np
t = np.array(buf, dtype=np.int64).reshape(16, 2, 32, 8)
nit = 16
names0 = ["top", "sdp_full", "kv_empty", "ld S,dP", "elementwise", "(last) store", "fence+arrive"]
names1 = ["top", "s_free", "issue S,dP", "ds_full", "issue dQ", "kv_empty"]
def med(a): return int(statistics.median(a))
print("compute thread 0 (cycles from previous stamp; median over CTAs and iterations 2..13)")
for sl in range(1, 7):
    print(f"  {names0[sl]:14s} {med([t[c,0,j,sl]-t[c,0,j,sl-1] for c in range(16) for j in range(2,14)])}")
print("  iteration period", med([t[c,0,j+1,0]-t[c,0,j,0] for c in range(16) for j in range(2,13)]))
print("control thread")
for sl in range(1, 6):
    print(f"  {names1[sl]:14s} {med([t[c,1,j,sl]-t[c,1,j,sl-1] for c in range(16) for j in range(2,13)])}")
print("  iteration period", med([t[c,1,j+1,0]-t[c,1,j,0] for c in range(16) for j in range(2,12)]))
c = 0
print("CTA 0, iterations 4..6 (cycles since iteration 4 top of compute):")
t0 = t[c,0,4,0]
for j in range(4, 7):
    print("  compute", j, [int(t[c,0,j,s]-t0) for s in range(7)], " control", [int(t[c,1,j,s]-t0) for s in range(6)])

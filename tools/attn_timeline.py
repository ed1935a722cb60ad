#!/usr/bin/env python
"""Bring-up aid: per-iteration timeline (SM clock cycles) of CTA 0 of the persistent attention kernel.
   role 0 = control thread:  0 loop top | 1 s_free seen | 2 S(g+1)+K load issued | 3 p_full seen | 4 PV issued | 5 V load issued
   role 1 = softmax thread0: 0 loop top | 1 s_full seen | 2 S in regs, s_free arrived | 3 exp+sum done | 4 P stored | 5 p_full arrived"""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from molly_b200 import _lib, ops

heads, d, k, n_seq = 20, 64, 1024, 64
h = heads * d
torch.manual_seed(0)
qkv = (torch.randn(n_seq * k, 3 * h, device="cuda") * 0.5).to(torch.bfloat16)
kv_info = torch.tensor([[k, k]] * n_seq, dtype=torch.int32, device="cuda")
mask = torch.ones(n_seq * k, dtype=torch.uint8, device="cuda")
lib = _lib.load()
for _ in range(3):
    ops.attention(qkv, n_seq, k, heads, kv_info, mask)
torch.cuda.synchronize()
buf = torch.zeros(2 * 64 * 8, dtype=torch.int64, device="cuda")
lib.molly_attention_debug(C.c_void_p(buf.data_ptr()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.attention(qkv, n_seq, k, heads, kv_info, mask)
e1.record()
torch.cuda.synchronize()
lib.molly_attention_debug(None)
t = buf.cpu().view(2, 64, 8)
print(f"kernel time {e0.elapsed_time(e1)*1e3:.1f} us; blocks/SM-pair-slot: {n_seq*heads*8*8/296:.1f}")
base = int(t[t > 0].min())
for role, name in ((0, "control"), (1, "softmax")):
    print(name)
    for g in range(0, 40):
        row = t[role, g, :6]
        if int(row.max()) == 0:
            break
        rel = [int(v) - base if v > 0 else -1 for v in row]
        d = [rel[i + 1] - rel[i] for i in range(5)]
        nxt = int(t[role, g + 1, 0]) - int(row[0]) if int(t[role, g + 1, 0]) > 0 else -1
        print(f"  g={g:2d} t0={rel[0]:7d} deltas={d} period={nxt}")

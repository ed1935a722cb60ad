#!/usr/bin/env python
"""Bring-up aid: clock64 timeline of the default attention kernel built with -DATT_TIMELINE (MOLLY_LIB=<that .so>;
run with MOLLY_ATTN_STREAM=0 so that one CTA = one item).
Per CTA: entry, setup done, per KV block (S seen, S in registers, exp+sum done, P handed over), O seen, stores done, exit;
control thread: loads issued, Q seen, S(0) issued, per block (s_free seen, p_full seen, PV drained)."""
import ctypes as C
import os
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from molly_b200 import _lib
if os.environ.get("MOLLY_LIB"):
    _lib.LIB_PATH = os.environ["MOLLY_LIB"]
from molly_b200 import ops

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
buf = torch.zeros(2048 * 72, dtype=torch.int64, device="cuda")
lib.molly_attention_debug(C.c_void_p(buf.data_ptr()))
ops.attention(qkv, n_seq, k, heads, kv_info, mask)
torch.cuda.synchronize()
lib.molly_attention_debug(None)
t = buf.cpu().numpy().reshape(2048, 72)
nkv = (k + 127) // 128 if os.environ.get("MOLLY_ATTN_KVB", "128") != "64" else 8   # only the first 8 blocks are recorded
ok = t[:, 36] > 0
t = t[ok]
print("CTAs recorded", len(t))
# steady-state CTAs only: skip the first wave (cold) -- take CTAs whose entry is later than the median exit of wave 0
life = t[:, 36] - t[:, 0]
print("CTA lifetime clk: median %d  p10 %d  p90 %d" % (np.median(life), np.percentile(life, 10), np.percentile(life, 90)))
def med(a): return int(np.median(a))
print("setup (entry -> after sync+TMEM alloc):", med(t[:, 1] - t[:, 0]))
print("control: setup -> loads issued:", med(t[:, 40] - t[:, 1]), " -> Q landed:", med(t[:, 41] - t[:, 40]),
      " -> S(0) issued:", med(t[:, 42] - t[:, 41]))
print("softmax: setup -> S(0) seen:", med(t[:, 2] - t[:, 1]))
# NOTE every timestamp costs a dependent global load of the buffer pointer (tens to hundreds of cycles): intervals are upper bounds
for j in range(6):
    b = 2 + 5 * j
    if not (t[:, b] > 0).all():
        break
    print(f"block {j}: wait-for-S {med(t[:, b] - (t[:, b - 1] if j else t[:, 1]))}  ld {med(t[:, b + 1] - t[:, b])}  "
          f"mask+max {med(t[:, b + 2] - t[:, b + 1])}  exp+sum {med(t[:, b + 3] - t[:, b + 2])}  "
          f"wait PV(j-1)+P store+arrive {med(t[:, b + 4] - t[:, b + 3])}  "
          f"| control: s_free->p_full {med(t[:, 44 + 3 * j] - t[:, 43 + 3 * j])}")
print("O seen -> stores done (first item of the CTA):", med(t[:, 35] - t[:, 34]), " final sync+dealloc:", med(t[:, 36] - t[:, 35]))
# per-SM gaps between a CTA's exit and the next CTA's entry on the same SM slot
sm = t[:, 37]
gaps = []
for s_ in np.unique(sm):
    rows = t[sm == s_]
    ent = np.sort(rows[:, 0]); ext = np.sort(rows[:, 36])
    for e in ent:
        prev = ext[ext <= e]
        if len(prev):
            gaps.append(e - prev.max())
if gaps:
    print("exit -> next entry on the same SM: median %d clk" % np.median(gaps))

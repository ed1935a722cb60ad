#!/usr/bin/env python
"""Bring-up aid: timeline of the ping-pong attention kernel (MOLLY_ATTN_PP=1, library built with -DATT_TIMELINE).
Per item and softmax group, per KV block: S seen | S in registers | token acquired | exp done | P handed over."""
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
t = t[300:2000]                     # steady-state items
def med(a): return int(np.median(a))
for x, name in ((0, "group A"), (1, "group B")):
    print(name)
    for j in range(7):
        b = 36 * x + 5 * j
        prev = t[:, b - 1] if j else None
        print(f"  block {j}: " + (f"prev handed->S seen {med(t[:, b] - prev):5d}  " if j else " " * 32) +
              f"ld {med(t[:, b + 1] - t[:, b]):4d}  mask+max+token wait {med(t[:, b + 2] - t[:, b + 1]):5d}  "
              f"exp {med(t[:, b + 3] - t[:, b + 2]):5d}  pack+store+handoff {med(t[:, b + 4] - t[:, b + 3]):5d}")
a_tok, b_tok = t[:, 2 + 5 * 2], t[:, 36 + 2 + 5 * 2]
print("A token(2) -> B token(2):", med(b_tok - a_tok), " B token(2) -> A token(3):", med(t[:, 2 + 5 * 3] - b_tok))
raw = buf.cpu().numpy().reshape(2048, 72)
per = [raw[i + 148, 0] - raw[i, 0] for i in range(300, 1800) if raw[i + 148, 0] > 0 and raw[i, 0] > 0]
print("item period per CTA (A block0 S seen -> next item's):", med(np.array(per)))
gap = [raw[i + 148, 0] - raw[i, 5 * 6 + 4] for i in range(300, 1800) if raw[i + 148, 0] > 0]
print("A: block 6 handed -> next item's block 0 S seen (contains block 7 + epilogue):", med(np.array(gap)))

#!/usr/bin/env python
"""Bring-up aid: clock64 timeline of attention2 built with the TL2 stamps (tools/build_variant.sh + patch; MOLLY_LIB=<that .so>).
Per (CTA < 8, tile, stream block g < 64): 0 block start, 1 S seen, 2 S in registers + s_free, 3 row max done,
4 first chunk's exp2 done, 5 PV(g-1) seen, 6 all P stored, 7 p_full arrived."""
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
buf = torch.zeros(8 * 2 * 64 * 8, dtype=torch.int64, device="cuda")
lib.molly_attention_debug(C.c_void_p(buf.data_ptr()))
ops.attention(qkv, n_seq, k, heads, kv_info, mask)
torch.cuda.synchronize()
lib.molly_attention_debug(None)
t = buf.cpu().numpy().reshape(8, 2, 64, 8)
np.save(os.environ.get("TL_OUT", "gpurun_out/attn2_timeline.npy"), t)
names = os.environ.get("TL_NAMES", "wait S,ld S,max,exp c0,wait PV,exp rest,arrive").split(",")
for cta in range(2):
    base = t[cta, :, 8, 0].min()
    for g in range(16, 28):
        row = []
        for tile in range(2):
            s = t[cta, tile, g]
            row.append(f"t{tile} start {s[0]-base:7d} | " + " ".join(f"{n} {s[i+1]-s[i]:5d}" for i, n in enumerate(names)) + f" | blk {t[cta,tile,g+1,0]-s[0]:5d}")
        print(f"cta{cta} g{g:2d}  " + "   ||   ".join(row))
d_ = np.diff(t[:, :, 16:56, :], axis=-1).reshape(-1, 7)
print("median phase lengths:", dict(zip(names, np.median(d_, axis=0).astype(int))))
per = (t[:, :, 17:57, 0] - t[:, :, 16:56, 0]).reshape(-1)
print("block period: median %d p10 %d p90 %d" % (np.median(per), np.percentile(per, 10), np.percentile(per, 90)))
# overlap of the exp phases (stamps 3..6) of the two tiles of one CTA
ov = []
for cta in range(8):
    a, b = t[cta, 0, 16:56], t[cta, 1, 16:56]
    for ga in range(40):
        e0, e1 = [int(x) for x in os.environ.get("TL_EXP", "3,6").split(",")]
        lo, hi = a[ga, e0], a[ga, e1]
        o = sum(max(0, min(hi, b[gb, e1]) - max(lo, b[gb, e0])) for gb in range(40))
        ov.append(o / max(1, hi - lo))
print("fraction of tile 0's exp phase during which tile 1 is also in its exp phase: median %.2f" % np.median(ov))

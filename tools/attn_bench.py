#!/usr/bin/env python
"""Isolated attention timing (median of 20 launches, CUDA events) for the two cfg-2 layer shapes.
   MOLLY_LIB=<path to an alternative libmolly_b200.so> selects the build under test."""
import os, sys, statistics
import torch
sys.path.insert(0, ".")
from molly_b200 import _lib
if os.environ.get("MOLLY_LIB"):
    _lib.LIB_PATH = os.environ["MOLLY_LIB"]
from molly_b200 import ops

def run(heads, d, k, n_seq, tag):
    h = heads * d
    torch.manual_seed(0)
    qkv = (torch.randn(n_seq * k, 3 * h, device="cuda") * 0.5).to(torch.bfloat16)
    kv_info = torch.tensor([[k, k]] * n_seq, dtype=torch.int32, device="cuda")
    mask = torch.ones(n_seq * k, dtype=torch.uint8, device="cuda")
    for _ in range(5):
        ops.attention(qkv, n_seq, k, heads, kv_info, mask)
    ts = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.attention(qkv, n_seq, k, heads, kv_info, mask); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = statistics.median(ts)
    fl = 4.0 * n_seq * k * k * h
    print(f"{tag}: {t*1e3:.1f} us  {fl/t/1e9:.1f} TFLOP/s")

print("lib", _lib.LIB_PATH, {k: v for k, v in os.environ.items() if k.startswith("MOLLY_ATTN")})
run(20, 64, 1024, 64, "ESM-650M layer (20 heads x 64, K=1024, 64 seqs)")
run(16, 64, 1024, 64, "NT-v2-500M layer (16 heads x 64, K=1024, 64 seqs)")

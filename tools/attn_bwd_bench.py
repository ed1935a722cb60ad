#!/usr/bin/env python
"""Isolated attention-backward timing (median of 20 launches, CUDA events): delta + dQ + dK/dV kernels of one layer.
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
    out, lse2 = ops.attention_lse(qkv, n_seq, k, heads, kv_info, mask)
    d_out = (torch.randn(n_seq * k, h, device="cuda") * 0.1).to(torch.bfloat16)
    for _ in range(5):
        ops.attention_bwd(qkv, out, d_out, lse2, n_seq, k, heads, kv_info, mask)
    ts = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.attention_bwd(qkv, out, d_out, lse2, n_seq, k, heads, kv_info, mask); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = statistics.median(ts)
    fl = 10.0 * n_seq * k * k * h
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            ops.attention_bwd(qkv, out, d_out, lse2, n_seq, k, heads, kv_info, mask)
        torch.cuda.synchronize()
    per = {}
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            nm = "dq" if "dq_kernel" in e.name else ("dkv" if "dkv_kernel" in e.name else ("delta" if "delta" in e.name else None))
            if nm:
                per.setdefault(nm, []).append(e.time_range.elapsed_us())
    split = ", ".join(f"{n} {statistics.median(v):.1f}" for n, v in per.items())
    print(f"{tag}: {t*1e3:.1f} us  {fl/t/1e9:.1f} TFLOP/s (2.5x the forward FLOP)   [{split} us]")

print("lib", _lib.LIB_PATH)
run(20, 64, 1024, 8, "ESM-650M layer, 8 seqs (train_bio)")
run(16, 64, 1024, 8, "NT-v2-500M layer, 8 seqs (train_bio)")
run(20, 64, 1024, 64, "ESM-650M layer, 64 seqs")

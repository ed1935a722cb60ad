#!/usr/bin/env python
"""Forward / backward split of the --train-bio step of one encoder (ESM-2 650M, 8 x 1024 tokens), both tape modes."""
import os, sys, time, statistics
import torch
sys.path.insert(0, ".")
import bench
from molly_b200 import train
from molly_b200.config import EncoderConfig
from molly_b200.packing import PackedEncoder

dev = torch.device("cuda", 0)
name = "esm2_t33_650m"
e = bench.ENC[name]
sd = bench.gpu_state_dict(e, dev, 1)
proj = {"weight": torch.zeros(64, e["hidden_size"]), "bias": torch.zeros(64)}
enc = PackedEncoder(EncoderConfig.from_mapping(dict(e, name=name)), sd, proj, 1024, dev)
wl = dict(bench.WORKLOADS["train_1p7b"])
ids = bench.make_inputs(wl)[0][:, 1].contiguous().to(dev)
d_out = (torch.randn(ids.numel(), e["hidden_size"], device=dev) * 0.01).to(torch.bfloat16)
for mode in ("1", "0"):
    os.environ["MOLLY_TRAIN_RECOMPUTE"] = mode
    fw, bw, wall = [], [], []
    for it in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        out, tape = train.encoder_forward_train(enc, ids)
        e1.record()
        grads = train.encoder_backward(enc, tape, d_out)
        e2.record()
        cpu_done = time.perf_counter() - t0
        torch.cuda.synchronize()
        if it >= 2:
            fw.append(e0.elapsed_time(e1)); bw.append(e1.elapsed_time(e2)); wall.append(cpu_done * 1e3)
        del grads, tape, out
    print(f"recompute={mode}: forward {statistics.median(fw):.1f} ms, backward {statistics.median(bw):.1f} ms, "
          f"host time to enqueue the step {statistics.median(wall):.1f} ms, peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")

if os.environ.get("TOP_KERNELS"):
    from torch.profiler import profile, ProfilerActivity
    os.environ["MOLLY_TRAIN_RECOMPUTE"] = "0"
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        out, tape = train.encoder_forward_train(enc, ids)
        grads = train.encoder_backward(enc, tape, d_out)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))

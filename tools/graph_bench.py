#!/usr/bin/env python
"""Latency of the fused input path (embed_tokens + encode + project + merge) at generate-sized batches: eager launches vs
one CUDA-graph replay.  Prints one JSON line per shape."""
import json
import statistics
import sys
import time
import torch
sys.path.insert(0, ".")
import bench
from molly_b200 import ops


def run(wl_name, B):
    wl = dict(bench.WORKLOADS[wl_name], B=B)
    dev = torch.device("cuda", 0)
    path = bench.build_path(wl, dev, strict=False)
    omic_ids, infos = bench.make_inputs(wl)
    input_ids = bench.build_input_ids(wl, infos)
    table = (torch.randn(bench.LLM_VOCAB, wl["D"], device=dev) * 0.02).to(torch.bfloat16)
    ids_dev, omic_dev = input_ids.to(dev), omic_ids.to(dev)
    types = [[i["type"] for i in row] for row in infos]
    call = path.graphed(table, B, wl["T"], types, wl["K"], bench.PAD_TOKEN_IDS)
    eager = lambda: path.embed_and_process(ids_dev, table, omic_dev, infos, bench.PAD_TOKEN_IDS)
    graph = lambda: call(ids_dev, omic_dev)
    assert torch.equal(eager(), graph())
    out = {"workload": wl["desc"], "B": B, "omics_tokens": B * 2 * wl["K"]}
    for name, fn in (("eager", eager), ("graph", graph)):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        gpu, wall = [], []
        for _ in range(30):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            wall.append((time.perf_counter() - t0) * 1e3)
            gpu.append(e0.elapsed_time(e1))
        l0 = ops.kernel_launch_count(); fn(); launches = ops.kernel_launch_count() - l0
        out[name] = {"gpu_ms": round(statistics.median(gpu), 4), "wall_ms": round(statistics.median(wall), 4),
                     "host_launches": launches}
    print(json.dumps(out), flush=True)
    path.close()


if __name__ == "__main__":
    for wl_name, B in (("molly_mini", 1), ("molly_mini", 4), ("molly_1p7b", 1), ("molly_1p7b", 4)):
        run(wl_name, B)

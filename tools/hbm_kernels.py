#!/usr/bin/env python
"""The HBM-bound kernels of the path at the headline (cfg-2) sizes: CUDA-event GB/s over algorithmic bytes, and -- under
`ncu --set full -k regex:...` (tools/gpu_hbm_ncu.sh) -- DRAM bytes / throughput per launch for profiles/r02_hbm_ncu.md.
Kernels: layernorm_kernel, embed_kernel, placeholder_runs_kernel, placeholder_reject_kernel, embed_tokens_skip_kernel,
merge_rows_kernel, placeholder_scan_kernel."""
import json
import statistics
import sys
import torch
sys.path.insert(0, ".")
import bench
from molly_b200 import _lib, ops
from molly_b200.config import EncoderConfig

dev = torch.device("cuda", 0)
wl = bench.WORKLOADS["molly_1p7b"]
B, K, T, D = wl["B"], wl["K"], wl["T"], wl["D"]
peak = bench.measured_peaks()["hbm_gbs"]


import os
N_TIMED, N_WARM = int(os.environ.get('HBM_N', 20)), int(os.environ.get('HBM_WARM', 3))


def timed(fn, n=None, warm=None, flush=None):
    n, warm = n or N_TIMED, N_WARM if warm is None else warm
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        if flush is not None:
            flush.fill_(1.0)                       # > 126 MB: the next launch starts from a cold L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev).view(torch.float32)
out = {}

def report(name, ms, nbytes):
    gbs = nbytes / (ms * 1e-3) / 1e9
    out[name] = {"ms": round(ms, 4), "algorithmic_bytes": int(nbytes), "gbs": round(gbs, 1), "frac_of_measured_hbm": round(gbs / peak, 3)}
    print(f"{name:28s} {ms*1e3:9.1f} us  {nbytes/1e6:9.1f} MB  {gbs:8.1f} GB/s  {gbs/peak:5.2f} of {peak:.0f}")

# LayerNorm fp32 -> bf16 on the ESM-650M residual stream: 65536 x 1280
for h in (1280, 1024):
    x = torch.randn(B * K, h, device=dev)
    w, b = torch.ones(h, device=dev), torch.zeros(h, device=dev)
    report(f"layernorm_kernel h={h}", timed(lambda: ops.layernorm(x, w, b, 1e-5), flush=flush), B * K * h * 6)
# embedding gather of one modality
for key in ("pr", "nt"):
    e = bench.ENC[wl[key]]
    cfg = EncoderConfig.from_mapping(dict(e, name=wl[key]))
    c = _lib.EncoderConfig(hidden_size=e["hidden_size"], num_layers=1, num_heads=e["num_attention_heads"],
                           intermediate_size=e["intermediate_size"], vocab_size=e["vocab_size"], pad_token_id=1,
                           mask_token_id=e["mask_token_id"], position_type=_lib.POS_ROTARY, max_positions=e["max_position_embeddings"],
                           ffn_type=0, token_dropout=int(e["token_dropout"]), emb_layer_norm_before=0,
                           layer_norm_eps=e["layer_norm_eps"], llm_hidden_size=D, project_token_num=K)
    ids, _ = bench.make_inputs(wl)
    idm = ids[:, 1 if key == "pr" else 0].contiguous().to(dev)
    table = (torch.randn(e["vocab_size"], e["hidden_size"], device=dev) * 0.02).to(torch.bfloat16)
    report(f"embed_kernel {wl[key]}", timed(lambda: ops.embed(idm, c, table, None), flush=flush),
           B * K * (8 + e["hidden_size"] * 4 + 1))          # ids read, fp32 row written, mask byte (table rows hit L2)
# input producer: run scan, reject pass, skipping LLM lookup
omic_ids, infos = bench.make_inputs(wl)
input_ids = bench.build_input_ids(wl, infos).to(dev)
table = (torch.randn(bench.LLM_VOCAB, D, device=dev) * 0.02).to(torch.bfloat16)
slots = torch.tensor([len(r) for r in infos], dtype=torch.int32, device=dev)
expect = torch.tensor([[1 if i["type"] == "protein" else 0 for i in r] for r in infos], dtype=torch.int32, device=dev)
runs = ops.placeholder_runs(input_ids, bench.PAD_TOKEN_IDS, slots, 2)
report("placeholder_runs_kernel", timed(lambda: ops.placeholder_runs(input_ids, bench.PAD_TOKEN_IDS, slots, 2), flush=flush), B * T * 12)
report("placeholder_reject_kernel", timed(lambda: ops.placeholder_reject(runs, expect, K, K)), B * 2 * 16)
text_rows = B * T - 2 * B * K
report("embed_tokens_skip_kernel", timed(lambda: ops.embed_tokens_skip(input_ids, runs[4], bench.PAD_TOKEN_IDS, K, K, table), flush=flush),
       B * T * 12 + 2 * text_rows * D * 2)
report("placeholder_scan_kernel", timed(lambda: ops.placeholder_scan(input_ids, *bench.PAD_TOKEN_IDS), flush=flush), B * T * 16)
# un-fused merge (the default path fuses it into the projector epilogue: 0 bytes)
hs = torch.zeros(B, T, D, dtype=torch.bfloat16, device=dev)
src = torch.randn(B * K, D, device=dev).to(torch.bfloat16)
seq_table = torch.tensor([[b, 20] for b in range(B)], dtype=torch.int32, device=dev)
report("merge_rows_kernel", timed(lambda: ops.merge_rows_(hs, src, seq_table, K, K), flush=flush), 2 * B * K * D * 2)
print(json.dumps(out))

#!/bin/bash
# ncu --set full on the two attention-backward kernels (ESM-2 650M layer shape, 8 sequences), plus the isolated timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/attn_bwd_bench.py 2>&1 | tail -4
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_d -c 2 -f -o gpurun_out/attn_bwd python tools/attn_bwd_bench.py > gpurun_out/attn_bwd_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/attn_bwd_ncu.log
ls -la gpurun_out/attn_bwd.ncu-rep

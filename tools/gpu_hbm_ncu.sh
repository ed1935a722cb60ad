#!/bin/bash
# CUDA-event GB/s + one ncu --set full capture of the HBM-bound kernels.
cd "$(dirname "$0")/.."
OUT=gpurun_out/hbm; mkdir -p $OUT
timeout -k 10 600 python tools/hbm_kernels.py > $OUT/events.txt 2>&1; echo "rc=$?"; cat $OUT/events.txt | tail -14
timeout -k 10 900 ncu --set full --clock-control none -k "regex:layernorm_kernel|embed_kernel|placeholder_runs_kernel|placeholder_reject_kernel|embed_tokens_skip_kernel|merge_rows_kernel|placeholder_scan_kernel" \
    -c 40 -o $OUT/prof_hbm -f env HBM_N=1 HBM_WARM=1 python tools/hbm_kernels.py > $OUT/ncu.log 2>&1; echo "ncu rc=$?"
ls -la $OUT

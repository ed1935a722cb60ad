#!/bin/bash
# attention backward: parity tests, isolated timing, and (when a -DBW_TIMELINE variant is present) the dQ kernel's timeline
cd "$(dirname "$0")/.."
timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py -q -x -p no:cacheprovider -k "attention" 2>&1 | tail -3
timeout 300 python tools/attn_bwd_bench.py 2>&1 | tail -4
for v in molly_b200/variants/libmolly_bwd_*.so; do
  [ -f "$v" ] || continue
  case $v in *bwd_tl.so) MOLLY_LIB=$v timeout 300 python tools/attn_bwd_timeline.py 2>&1 | tail -22; continue;; esac
  MOLLY_LIB=$v timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py -q -x -p no:cacheprovider -k "attention_backward" 2>&1 | tail -1
  MOLLY_LIB=$v timeout 300 python tools/attn_bwd_bench.py 2>&1 | tail -4
done

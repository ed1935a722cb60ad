#!/bin/bash
# attention parity + isolated timing (MOLLY_ATTN_V2 = 0 | 1, optional POLY)
cd "$(dirname "$0")/.."
timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "attention and not backward" -p no:cacheprovider 2>&1 | tail -6
for v in "MOLLY_ATTN_V2=0" "MOLLY_ATTN_V2=1" "MOLLY_ATTN_V2=1 MOLLY_ATTN_POLY=1" "MOLLY_ATTN_V2=1 MOLLY_ATTN_POLY=2"; do
  echo "-- $v"; env $v timeout -k 10 300 python tools/attn_bench.py 2>&1 | tail -2
done

#!/bin/bash
# A/B timing of attention variants built by tools/build_variant.sh (molly_b200/variants/*.so), isolated launches.
cd "$(dirname "$0")/.."
OUT=gpurun_out/${TAG:-r02c}; mkdir -p $OUT
for lib in ${LIBS:-molly_b200/variants/*.so}; do
  for v in ${VARS:-"MOLLY_ATTN_V2=1"}; do
    echo "-- $lib $v"; env $v MOLLY_LIB=$lib timeout -k 10 300 python tools/attn_bench.py 2>&1 | tail -2
  done
done

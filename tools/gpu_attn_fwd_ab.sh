#!/bin/bash
# forward attention: parity tests, then isolated timing of the default library and of a variant (MOLLY_LIB)
cd "$(dirname "$0")/.."
timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py -q -x -p no:cacheprovider -k "attention and not backward" 2>&1 | tail -2
for lib in "" "$@"; do
  echo "== ${lib:-default}"
  MOLLY_LIB=$lib timeout 300 python tools/attn_bench.py 2>&1 | tail -2
done

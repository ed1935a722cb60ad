#!/bin/bash
# A/B of the headline step on ONE box: default library vs a variant (MOLLY_LIB), alternating, 2 runs each
cd "$(dirname "$0")/.."
V=${1:?variant .so}; OUT=gpurun_out/ab; mkdir -p $OUT
for i in 1 2; do
  for lib in "" "$V"; do
    MOLLY_LIB=$lib timeout -k 10 600 python bench.py --headline-only --no-cpu-baseline --steps 8 --warmup 3 > $OUT/b.json 2> $OUT/b.err
    python - <<PY
import json
d = json.loads(open("$OUT/b.json").read().strip().splitlines()[-1])
k = d["kernels"]
print("${lib:-default}".split("/")[-1], "ms", round(d["ms_per_step"], 2), {n: k[n]["ms"] for n in ("gemm_qkv", "attention", "gemm_attn_out", "gemm_ffn1", "gemm_ffn2")}, d["clocks"]["sm_mhz"])
PY
  done
done

#!/bin/bash
# A/B of the headline step on ONE box: default vs an environment switch (e.g. MOLLY_RESID_REDUCE=1), alternating, 2 runs each
cd "$(dirname "$0")/.."
SW=${1:?VAR=value}; OUT=gpurun_out/ab; mkdir -p $OUT
env $SW timeout -k 10 600 python -m pytest tests/test_gpu_path.py tests/test_gpu_configs.py -q -x -p no:cacheprovider -k "golden or config" 2>&1 | tail -2
for i in 1 2; do
  for sw in "MOLLY_AB_NONE=1" "$SW"; do
    env $sw timeout -k 10 600 python bench.py --headline-only --no-cpu-baseline --steps 8 --warmup 3 > $OUT/b.json 2> $OUT/b.err
    python - <<PY
import json
d = json.loads(open("$OUT/b.json").read().strip().splitlines()[-1])
k = d["kernels"]
print("$sw", "ms", round(d["ms_per_step"], 2), {n: k[n]["ms"] for n in ("gemm_qkv", "attention", "gemm_attn_out", "gemm_ffn1", "gemm_ffn2", "layernorm")}, d["clocks"]["sm_mhz"])
PY
  done
done

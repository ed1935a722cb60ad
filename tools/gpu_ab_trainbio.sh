#!/bin/bash
# A/B of the --train-bio step on ONE box over environment switches: tools/gpu_ab_trainbio.sh "VAR=a" "VAR=b" ...
cd "$(dirname "$0")/.."
OUT=gpurun_out/abt; mkdir -p $OUT
for sw in "$@"; do
  env $sw timeout -k 10 600 python bench.py --workload train_bio_1p7b --steps 5 --warmup 3 > $OUT/b.json 2> $OUT/b.err
  python - <<PY
import json
d = json.loads(open("$OUT/b.json").read().strip().splitlines()[-1])
k = d["kernels"]
print("$sw", "ms", round(d["ms_per_step"], 2), {n: k[n]["ms"] for n in ("gemm_other", "rowwise_bwd", "attention_bwd")}, d["clocks"]["sm_mhz"])
PY
done

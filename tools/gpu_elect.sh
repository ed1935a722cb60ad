#!/bin/bash
# elect.sync control threads: parity tests, then isolated attention forward / backward timing against the lane-test build
cd "$(dirname "$0")/.."
OUT=gpurun_out/elect; mkdir -p $OUT
timeout -k 10 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_train.py -q -x -p no:cacheprovider > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
for lib in "" molly_b200/variants/libmolly_noelect.so; do
  echo "== lib=${lib:-default}"
  MOLLY_LIB=$lib timeout 300 python tools/attn_bench.py 2>&1 | tail -2
  MOLLY_ATTN_V2=1 MOLLY_LIB=$lib timeout 300 python tools/attn_bench.py 2>&1 | tail -2
  MOLLY_LIB=$lib timeout 300 python tools/attn_bwd_bench.py 2>&1 | tail -3
done
timeout -k 10 600 python bench.py --headline-only --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 2))
print({k: (v.get("ms"), v.get("tflops", v.get("gbs"))) for k, v in d["kernels"].items()})
print(d["clocks"])
PY

#!/bin/bash
# Round-2 bring-up of attention2 (one CTA per SM, two tiles, P in TMEM): parity, isolated A/B timing, ncu, then the whole suite.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${TAG:-r02b}
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
echo "== attention parity (v2 default)"
timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "attention and not backward" -p no:cacheprovider 2>&1 | tail -15
echo "== isolated timing"
for v in "MOLLY_ATTN_V2=0" "MOLLY_ATTN_V2=1" "MOLLY_ATTN_V2=1 MOLLY_ATTN_POLY=1" "MOLLY_ATTN_V2=1 MOLLY_ATTN_POLY=2"; do
  echo "-- $v"; env $v timeout -k 10 300 python tools/attn_bench.py 2>&1 | tail -3
done
echo "== ncu attention2"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:attention2 -s 5 -c 1 \
    -o $OUT/prof_attn2 -f python tools/attn_bench.py > $OUT/ncu_attn2.log 2>&1; echo "rc=$?"
echo "== full gpu suite"
timeout -k 10 1200 python -m pytest tests -q -x -m gpu -p no:cacheprovider 2>&1 | tail -15
echo "== bench"
timeout -k 10 600 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02b/bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
print({k: (v.get("ms"), v.get("tflops", v.get("gbs"))) for k, v in d["kernels"].items()})
print(d["clocks"])
PY

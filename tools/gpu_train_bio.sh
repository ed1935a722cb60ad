#!/bin/bash
# --train-bio step: bench line (kernel families, host enqueue time) + per-kernel times from a torch profiler pass.
# Usage: tools/gpu_train_bio.sh [tag]
cd "$(dirname "$0")/.."
TAG=${1:-r02}; OUT=gpurun_out/trainbio_$TAG; mkdir -p $OUT
timeout -k 10 600 python bench.py --workload train_bio_1p7b --steps ${STEPS:-5} --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -3 $OUT/bench.err
python - <<PY
import json
d = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print("ms", round(d["ms_per_step"], 2), "host enqueue", d.get("host_enqueue_ms_per_step"), "launches/step", d["gpu_launches"] / d["steps"])
tot = 0
for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms"]):
    print(f"  {k:16s} {v['launches']:5d} {v['ms']:8.3f}"); tot += v["ms"]
print("  sum", round(tot, 2), d["clocks"])
PY
if [ "${TRACE:-1}" = "1" ]; then timeout -k 10 600 python tools/train_bio_trace.py > $OUT/trace.txt 2>&1; echo "trace rc=$?"; head -70 $OUT/trace.txt; fi

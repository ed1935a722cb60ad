#!/bin/bash
# Full GPU check: the gpu test suite, smoke, then the default bench line (all sections).   Usage: tools/gpu_full.sh [tag]
cd "$(dirname "$0")/.."
TAG=${1:-r02}; OUT=gpurun_out/full_$TAG; mkdir -p $OUT
timeout -k 10 1500 python -m pytest tests -q -x -m gpu -p no:cacheprovider > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest.log
grep -E "pass-rate|real OmicsOne|NT-v2 variant" $OUT/pytest.log | head -20
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout -k 10 900 python bench.py --steps ${STEPS:-10} --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -3 $OUT/bench.err
python - <<PY
import json
d = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), "resident", round(d["e2e_device_resident"]["value"]))
print({k: (v.get("ms"), v.get("tflops", v.get("gbs"))) for k, v in d["kernels"].items()})
print("roofline", d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["frac_of_nominal"])
print("lib", json.dumps(d.get("gpu_library_baseline"))[:1500])
print("train", json.dumps({k: v for k, v in d.get("train_step", {}).items() if k != "kernels"})[:1200])
print("varlen", json.dumps(d.get("varlen"))[:800])
print("cpu", d.get("cpu_baseline"))
print(d["clocks"])
PY

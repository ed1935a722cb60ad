#!/bin/bash
# N GPUs of one box: headline line (--headline-only) and the --train-bio line.   Usage (gpurun --gpus N): tools/gpu_n8.sh N [tag]
cd "$(dirname "$0")/.."
N=${1:-8}; TAG=${2:-r02}; OUT=gpurun_out/n${N}_$TAG; mkdir -p $OUT
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --headline-only --no-cpu-baseline --steps 8 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "headline rc=$?"
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --workload train_bio_1p7b --steps 5 --warmup 3 > $OUT/bench_train_bio.json 2> $OUT/bench_train_bio.err; echo "train_bio rc=$?"
python - <<PY
import json
for f in ("bench.json", "bench_train_bio.json"):
    try:
        d = json.loads(open("$OUT/" + f).read().strip().splitlines()[-1])
        print(f, "n_gpus", d["n_gpus"], "ms", round(d["ms_per_step"], 2), "tok/s", round(d["value"]), "e2e", round(d.get("e2e", {}).get("value", 0)), json.dumps(d.get("comm"))[:300], d["clocks"])
    except Exception as e:
        print(f, "failed", e)
PY

#!/bin/bash
# 2 GPUs: averaged-gradient check of the layer-wise reducer (incl. a rank without one modality), then the --train-bio and
# --train-mlp (cfg-5) bench lines at N=2.   Usage (gpurun --gpus 2): tools/gpu_train_2gpu.sh [tag]
cd "$(dirname "$0")/.."
TAG=${1:-r02}; OUT=gpurun_out/train2_$TAG; mkdir -p $OUT
P=29541
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P tools/check_reducer_2gpu.py > $OUT/check_reducer.txt 2>&1; echo "check rc=$?"; tail -12 $OUT/check_reducer.txt
for wl in train_bio_1p7b train_1p7b; do
  timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus 2 --workload $wl --steps 5 --warmup 3 > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err; echo "$wl rc=$?"
  python - <<PY
import json
d = json.loads(open("$OUT/bench_$wl.json").read().strip().splitlines()[-1])
print("$wl", "ms", round(d["ms_per_step"], 2), "tok/s", round(d["value"]), "comm", json.dumps(d.get("comm"))[:600])
PY
done

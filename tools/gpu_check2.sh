#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/check2
for i in 1 2; do
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$i tools/check_reducer_2gpu.py > gpurun_out/check2/run$i.txt 2>&1; echo "check rc=$?"; grep "^rank" gpurun_out/check2/run$i.txt | head -40
done

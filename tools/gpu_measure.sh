#!/bin/bash
# Bench + profiling pass on the GPU box: bench line, ncu launch list, ncu --set full of the top kernels.
# Outputs land in gpurun_out/ (copied into profiles/ by hand once read).   Usage: tools/gpu_measure.sh [tag]
set -u
cd "$(dirname "$0")/.."
TAG=${1:-r01}
OUT=gpurun_out/measure_$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> $OUT/gpu.txt
echo "== bench (full default run)"
timeout -k 10 900 python bench.py --steps ${STEPS:-5} --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"
tail -c 3000 $OUT/bench.json; tail -5 $OUT/bench.err
if [ "${NCU:-1}" = "1" ]; then
  echo "== ncu launch list (one step, serialised, cold-cache: compare SHARES)"
  timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none \
      -k "regex:gemm_tcgen05_kernel|attention|layernorm_kernel|embed_kernel|rotary_kernel|mask_rows_kernel" \
      -s ${NCU_SKIP:-1506} -c ${NCU_COUNT:-502} --csv \
      --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --headline-only > $OUT/ncu_launch.log 2>&1
  echo "rc=$? lines=$(wc -l < $OUT/launches.csv)"
  echo "== ncu --set full: GEMMs of one encoder layer + attention"
  timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 8 -c 4 \
      -o $OUT/prof_gemm -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --headline-only > $OUT/ncu_gemm.log 2>&1; echo "rc=$?"
  timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:attention -s 2 -c 1 \
      -o $OUT/prof_attn -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --headline-only > $OUT/ncu_attn.log 2>&1; echo "rc=$?"
  ls -la $OUT
fi

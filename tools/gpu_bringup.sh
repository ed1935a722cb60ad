#!/bin/bash
# Runs every GPU test group in its OWN process (a trapped kernel poisons only its group), with timeouts,
# and collects logs + a summary under gpurun_out/bringup/.   Usage: tools/gpu_bringup.sh [group-filter-regex]
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/bringup
mkdir -p $OUT
FILTER=${1:-.}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > $OUT/gpu.txt 2>&1
run() {  # name timeout cmd...
  local name=$1 to=$2; shift 2
  if ! [[ $name =~ $FILTER ]]; then return; fi
  local t0=$(date +%s)
  timeout -k 10 $to "$@" > $OUT/$name.log 2>&1
  local rc=$?
  echo "$name rc=$rc secs=$(( $(date +%s) - t0 )) :: $(grep -E 'passed|failed|error' $OUT/$name.log | tail -1)" | tee -a $OUT/summary.txt
}
: > $OUT/summary.txt
PT="python -m pytest -q --tb=short -rA -p no:cacheprovider"
run simple        400 $PT tests/test_gpu_kernels.py -k "layernorm or rotary or embed or placeholder or merge_rows"
run gemm_bias     400 $PT tests/test_gpu_kernels.py -k "gemm_bias"
run gemm_misc     400 $PT tests/test_gpu_kernels.py -k "gemm_nobias or gemm_gelu or gemm_residual or gemm_glu or gemm_scatter or gemm_qkv_rope"
run attn_d64      400 $PT tests/test_gpu_kernels.py -k "test_attention and 64-"
run attn_other    400 $PT tests/test_gpu_kernels.py -k "test_attention and not 64-"
run path_encoder  600 $PT tests/test_gpu_path.py -k "encoder_forward"
run path_golden   600 $PT tests/test_gpu_path.py -k "golden"
run path_rest     900 $PT tests/test_gpu_path.py -k "not golden and not encoder_forward"
run backward      600 $PT tests/test_gpu_kernels.py -k "backward or linear_wgrad or log_sum_exp"
run inputs_graph  600 $PT tests/test_gpu_inputs.py tests/test_gpu_graph.py
run train         900 $PT tests/test_gpu_train.py
run configs       1200 $PT tests/test_gpu_configs.py
run variants      1200 $PT tests/test_gpu_variants.py
run smoke         300 python __graft_entry__.py smoke
echo "---- summary ----"; cat $OUT/summary.txt
# compact failure digest (full logs stay in gpurun_out/bringup/)
for f in $OUT/*.log; do
  if grep -qE "failed|rror" $f; then echo "=== $f"; grep -hE "^(E  |FAILED|molly:|.*Error)" $f | cut -c1-220 | head -${DIGEST_LINES:-40}; fi
done | head -300

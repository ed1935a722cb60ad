#!/bin/bash
# Build deliberately-wrong ablated variants of the attention kernel (one piece removed each) next to the real library and
# time them with tools/attn_bench.py: what each piece costs.  Run here to build (nvcc), then on the GPU box:
#   tools/attn_ablate.sh build            (container)
#   tools/attn_ablate.sh run              (GPU box: prints one line per variant and KV block size)
set -e
cd "$(dirname "$0")/.."
CS=molly_b200/csrc
OUT=molly_b200/_ablate
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
MASKS="0 1 2 3 4 8 11 15 16 20 32 63"
if [ "$1" = build ]; then
  mkdir -p $OUT
  for m in $MASKS; do
    ( nvcc $FLAGS -DATT_ABLATE=$m -I$CS -c $CS/attention.cu -o $OUT/attention_$m.o &&
      nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libmolly_abl_$m.so $CS/build/common.o $CS/build/gemm.o \
        $OUT/attention_$m.o $CS/build/rowwise.o $CS/build/merge.o $CS/build/bwd.o $CS/build/abi.o && rm $OUT/attention_$m.o ) &
  done
  wait
  ls $OUT
else
  for kvb in 128 64; do for m in $MASKS; do
    echo -n "KVB=$kvb ablate=$m: "
    MOLLY_ATTN_KVB=$kvb MOLLY_LIB=$PWD/$OUT/libmolly_abl_$m.so python tools/attn_bench.py 2>&1 | grep ESM | sed 's/.*: //'
  done; done
fi

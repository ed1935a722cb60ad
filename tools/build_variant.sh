#!/bin/bash
# Build a variant libmolly_b200.so into molly_b200/variants/ (git-ignored, shipped by gpurun): tools/build_variant.sh <tag> [git-rev-of-attention2.cu] [EXTRA nvcc flags]
set -e
cd "$(dirname "$0")/.."
TAG=$1; REV=${2:-}; EXTRA_FLAGS=${3:-}
W=/tmp/molly_variant_$TAG
rm -rf $W; mkdir -p $W/molly_b200 $W/include
cp -r molly_b200/csrc $W/molly_b200/csrc; cp include/*.h $W/include/
rm -rf $W/molly_b200/csrc/build
if [ -n "$REV" ] && [ "$REV" != "-" ]; then git show $REV:molly_b200/csrc/attention2.cu > $W/molly_b200/csrc/attention2.cu; fi
make -C $W/molly_b200/csrc -j8 EXTRA="$EXTRA_FLAGS" > $W/build.log 2>&1 || (tail -20 $W/build.log; exit 1)
mkdir -p molly_b200/variants
cp $W/molly_b200/libmolly_b200.so molly_b200/variants/libmolly_$TAG.so
grep -A2 "attention2_kernelILi64ELi0" $W/molly_b200/csrc/build/attention2.ptxas.log | grep -E "spill|Used" | head -2
echo "built molly_b200/variants/libmolly_$TAG.so"

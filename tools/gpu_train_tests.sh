#!/bin/bash
# Training-path GPU tests (both tape modes) + the --train-bio bench line.   Usage: tools/gpu_train_tests.sh [tag]
cd "$(dirname "$0")/.."
TAG=${1:-r02}; OUT=gpurun_out/traintest_$TAG; mkdir -p $OUT
timeout -k 10 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_path.py tests/test_gpu_real_class.py tests/test_gpu_kernels.py -q -x -p no:cacheprovider > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest.log
MOLLY_TRAIN_RECOMPUTE=1 timeout -k 10 900 python -m pytest tests/test_gpu_train.py -q -x -p no:cacheprovider -k backward > $OUT/pytest_recompute.log 2>&1; echo "pytest recompute rc=$?"; tail -3 $OUT/pytest_recompute.log
TRACE=${TRACE:-0} bash tools/gpu_train_bio.sh $TAG

#!/bin/bash
# compute-sanitizer memcheck over the kernel tests and the small training-path tests.  Besides out-of-bounds accesses this is a
# protocol test: the instrumentation spreads the threads of a CTA far apart in time, so an mbarrier hand-over that silently
# relies on the threads staying within one block of each other shows up as a parity failure.
cd "$(dirname "$0")/.."
OUT=gpurun_out/sanitize; mkdir -p $OUT
timeout -k 10 2400 compute-sanitizer --tool memcheck --error-exitcode 99 --log-file $OUT/memcheck.log python -m pytest tests/test_gpu_kernels.py tests/test_gpu_train.py -q -p no:cacheprovider -k "not full_size" > $OUT/memcheck_pytest.log 2>&1; echo "memcheck rc=$?"; grep "passed\|failed" $OUT/memcheck_pytest.log | tail -3; grep "^FAILED" $OUT/memcheck_pytest.log | head -20; tail -3 $OUT/memcheck.log
if [ "${MORE:-0}" = "1" ]; then
  MOLLY_ATTN_V2=1 timeout -k 10 1200 compute-sanitizer --tool memcheck --error-exitcode 99 --log-file $OUT/memcheck_v2.log python -m pytest tests/test_gpu_kernels.py -q -p no:cacheprovider -k "attention and not backward" > $OUT/memcheck_v2_pytest.log 2>&1; echo "attention2 memcheck rc=$?"; grep "passed\|failed" $OUT/memcheck_v2_pytest.log | tail -2; tail -2 $OUT/memcheck_v2.log
  timeout -k 10 2400 compute-sanitizer --tool memcheck --error-exitcode 99 --log-file $OUT/memcheck_path.log python -m pytest tests/test_gpu_path.py tests/test_gpu_inputs.py tests/test_gpu_graph.py -q -p no:cacheprovider > $OUT/memcheck_path_pytest.log 2>&1; echo "path memcheck rc=$?"; grep "passed\|failed" $OUT/memcheck_path_pytest.log | tail -2; grep "^FAILED" $OUT/memcheck_path_pytest.log | head; tail -2 $OUT/memcheck_path.log
fi

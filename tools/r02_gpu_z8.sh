#!/bin/bash
# compute-sanitizer memcheck over the kernels touched at the end of the round: the input-pipeline kernel (4 x 4 work items, 32-bit
# index arithmetic) and the test-time plan with relu / staging absorbed into batchNormInference
set -u
TAG=${1:-r02z8}
OUT=gpurun_out
mkdir -p $OUT
SAN=/usr/local/cuda/bin/compute-sanitizer
timeout -k 5 45 $SAN --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest -x -q -p no:cacheprovider tests/test_input_pipeline.py > $OUT/${TAG}_sanitizer_memcheck_input.log 2>&1
echo "exit code $?" >> $OUT/${TAG}_sanitizer_memcheck_input.log
grep -E "ERROR SUMMARY|passed|failed|exit code" $OUT/${TAG}_sanitizer_memcheck_input.log | tail -4
timeout -k 5 45 $SAN --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest -x -q -p no:cacheprovider tests/test_baseline_configs_gpu.py -k "test_time" > $OUT/${TAG}_sanitizer_memcheck_infer.log 2>&1
echo "exit code $?" >> $OUT/${TAG}_sanitizer_memcheck_infer.log
grep -E "ERROR SUMMARY|passed|failed|exit code" $OUT/${TAG}_sanitizer_memcheck_infer.log | tail -4

#!/bin/bash
# compute-sanitizer over the per-op kernels and one small plan (run on the GPU box): memcheck on the convolution / batch-norm /
# matmul / pointwise tests and the small-convnet plan tests, racecheck + synccheck on the tensor-core convolution cases.
# Summaries land in gpurun_out/<tag>_sanitizer_*.log; copy them to profiles/.
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
SAN=/usr/local/cuda/bin/compute-sanitizer
run() {   # name, tool, pytest selection...
  local name=$1 tool=$2; shift 2
  timeout 1200 $SAN --tool $tool --error-exitcode 99 --print-limit 20 python -m pytest -x -q -p no:cacheprovider "$@" > $OUT/${TAG}_sanitizer_${name}.log 2>&1
  echo "exit code $?" >> $OUT/${TAG}_sanitizer_${name}.log
  grep -E "ERROR SUMMARY|passed|failed|exit code" $OUT/${TAG}_sanitizer_${name}.log | tail -4
}
run memcheck_ops memcheck tests/test_ops_gpu.py -k "convolution_family or convolution_reference or batchnorm_train or matmul_tensor_core or pointwise_unary or relu_and_grad or softmax or maxpool_and_grad"
run memcheck_plan memcheck tests/test_plan_gpu.py -k "small_convnet_sgd_bf16 or wrn_16_4_sgd_bf16_interior or small_convnet_other"
run racecheck_conv racecheck tests/test_ops_gpu.py -k "convolution_family or matmul_tensor_core"
run synccheck_conv synccheck tests/test_ops_gpu.py -k "convolution_family or batchnorm_train"

#!/bin/bash
# Sweep of the batch-norm statistics grid (round-2 experiment; run on the GPU box from the repo root):
#   gpurun --timeout 300 -- 'bash tools/bn_sweep.sh > gpurun_out/bn_sweep.txt 2>&1'
# Prints images/s, ms/step and the device time of the two batch-norm op types for each setting.
for v in "" "DOPT_B200_BN_CTAS_PER_SM=5" "DOPT_B200_BN_CTAS_PER_SM=6" "DOPT_B200_BN_CTAS_PER_SM=8" "DOPT_B200_BN_CTAS_PER_SM=12"; do
  env $v timeout 60 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.readline())
p = d['per_op_us_per_step']
print('%-32s %8.0f img/s %7.3f ms  bnTrain %7.1f us  bnGrad %7.1f us' % ('$v' or 'default', d['value'], d['ms_per_step'], p['batchNormTrain'][0], p['batchNormGrad'][0]))
"
done

#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_plan_gpu.py -q -k "interior or 28_10 or wrn_16" -s > $OUT/r02d_pytest_new.log 2>&1
echo "pytest new rc=$?" >> $OUT/r02d_pytest_new.log
grep -E "update errors|losses|passed|failed|Error|error" $OUT/r02d_pytest_new.log | tail -30
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flat_bn_stats --launch-skip 60 --launch-count 3 -o $OUT/r02d_flat_stats python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> $OUT/r02d_ncu.err
tail -3 $OUT/r02d_ncu.err
ls -la $OUT/*.ncu-rep

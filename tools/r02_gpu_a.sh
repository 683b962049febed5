#!/bin/bash
# round-2 GPU call A: plan-level parity of the bf16-interior path + a first bench / per-op profile / timeline
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_plan_gpu.py -q -x -k "interior or 28_10 or wrn_16" -s > $OUT/r02a_pytest_new.log 2>&1
echo "pytest new rc=$?" >> $OUT/r02a_pytest_new.log
tail -30 $OUT/r02a_pytest_new.log
DOPT_B200_PLAN_DUMP=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/r02a_bench.json 2> $OUT/r02a_bench.err
grep "PLAN residency" $OUT/r02a_bench.err
cat $OUT/r02a_bench.json
timeout 200 python bench.py --timeline $OUT/r02a_timeline.txt --no-cpu-baseline > /dev/null 2>> $OUT/r02a_bench.err
head -40 $OUT/r02a_timeline.txt

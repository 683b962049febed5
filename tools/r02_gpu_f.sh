#!/bin/bash
set -u
OUT=gpurun_out
TAG=${1:-r02f}
mkdir -p $OUT
DOPT_B200_PLAN_DUMP=1 timeout 200 python bench.py --timeline $OUT/${TAG}_timeline.txt --no-cpu-baseline > /dev/null 2> $OUT/${TAG}.err
grep "PLAN residency" $OUT/${TAG}.err
head -18 $OUT/${TAG}_timeline.txt
timeout 900 python -m pytest tests/test_plan_gpu.py -q -k "interior or 28_10 or wrn_16" -s > $OUT/${TAG}_pytest_new.log 2>&1
echo "pytest new rc=$?" >> $OUT/${TAG}_pytest_new.log
grep -E "passed|failed|Error|error" $OUT/${TAG}_pytest_new.log | tail -10
grep -E "update errors|losses" $OUT/${TAG}_pytest_new.log | cut -c1-400
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2>> $OUT/${TAG}.err
python -c "
import json; d=json.load(open('$OUT/${TAG}_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], {k:(round(v['frac'],3), v['us_per_step']) for k,v in d['roofline_classes'].items()}); print(d['per_op_us_per_step']); print(d['loss_first'], d['loss_last'])"

#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 200 python bench.py --timeline $OUT/r02c_timeline.txt --no-cpu-baseline > /dev/null 2> $OUT/r02c.err
head -16 $OUT/r02c_timeline.txt; tail -5 $OUT/r02c.err
for c in 1 4; do
  DOPT_B200_FLAT_CLUSTER=$c timeout 200 python bench.py --timeline $OUT/r02c_timeline_cl$c.txt --no-cpu-baseline > /dev/null 2>> $OUT/r02c.err
  echo "== FLAT_CLUSTER=$c"; grep -E "busy|flat_" $OUT/r02c_timeline_cl$c.txt
done
timeout 900 python -m pytest tests/test_plan_gpu.py -q -k "interior or 28_10 or wrn_16" -s > $OUT/r02c_pytest_new.log 2>&1
echo "pytest new rc=$?" >> $OUT/r02c_pytest_new.log
grep -E "update errors|losses|passed|failed|Error|error" $OUT/r02c_pytest_new.log | tail -30
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/r02c_bench.json 2>> $OUT/r02c.err
python -c "
import json; d=json.load(open('$OUT/r02c_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], {k:(round(v['frac'],3), v['us_per_step']) for k,v in d['roofline_classes'].items()}); print(d['per_op_us_per_step'])"

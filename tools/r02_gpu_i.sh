#!/bin/bash
set -u
OUT=gpurun_out
TAG=${1:-r02i}
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_plan_gpu.py -q -x > $OUT/${TAG}_pytest_plan.log 2>&1
echo "pytest plan rc=$?" >> $OUT/${TAG}_pytest_plan.log
tail -5 $OUT/${TAG}_pytest_plan.log
timeout 200 python bench.py --timeline $OUT/${TAG}_timeline.txt --no-cpu-baseline > /dev/null 2> $OUT/${TAG}.err
head -22 $OUT/${TAG}_timeline.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2>> $OUT/${TAG}.err
python -c "
import json; d=json.load(open('$OUT/${TAG}_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], {k:(round(v['frac'],3), v['us_per_step']) for k,v in d['roofline_classes'].items()}); print(d['per_op_us_per_step']); print(d['loss_first'], d['loss_last'], d['launches_per_step'])"

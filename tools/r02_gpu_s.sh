#!/bin/bash
# round-2 GPU call S: flat batch-norm apply kernel with every prologue load issued before the first use
set -u
OUT=gpurun_out
TAG=${1:-r02s}
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_plan_gpu.py -q -x -k "28_10 or wrn_16 or interior or regular" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
tail -4 $OUT/${TAG}_pytest.log
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_$i.json 2> $OUT/${TAG}_$i.err
python - <<PY
import json
f = "$OUT/${TAG}_bench_$i.json"
try:
    d = json.load(open(f)); print("run$i", round(d['value']), round(d['ms_per_step'], 3), round(d['e2e']['value']), d['roofline']['kernel_ms_per_step'], round(d['roofline']['frac_of_burst_peak'], 3), d['loss_first'], d['loss_last'], d['launches_per_step'], {k:(round(v['frac'],3), round(v['us_per_step'])) for k,v in d['roofline_classes'].items()})
except Exception as e: print(f, "FAILED", e)
PY
done
DOPT_B200_NO_SIDE_STREAM=1 DOPT_B200_PDL=0 timeout 200 python bench.py --timeline $OUT/${TAG}_timeline_serial.txt --no-cpu-baseline > /dev/null 2>> $OUT/${TAG}.err
head -16 $OUT/${TAG}_timeline_serial.txt

#!/bin/bash
# round-2 GPU call N: filter-gradient halo variant (three vertical taps per item) -- parity, A/B bench, timelines
set -u
OUT=gpurun_out
TAG=${1:-r02n}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_cudnn_replay_gpu.py tests/test_ops_gpu.py -q -x -k "conv" > $OUT/${TAG}_pytest_conv.log 2>&1
echo "pytest conv rc=$?" >> $OUT/${TAG}_pytest_conv.log
tail -6 $OUT/${TAG}_pytest_conv.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}.err
DOPT_B200_WG_HALO=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_wg0.json 2>> $OUT/${TAG}.err
python - <<PY
import json
for f in ("$OUT/${TAG}_bench.json", "$OUT/${TAG}_bench_wg0.json"):
    try:
        d = json.load(open(f)); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms_per_step'], d['roofline']['frac_of_burst_peak'], d['loss_first'], d['loss_last'], d['per_op_us_per_step'].get('convolutionFiltersGrad'))
    except Exception as e: print(f, "FAILED", e)
PY
tail -5 $OUT/${TAG}.err
DOPT_B200_NO_SIDE_STREAM=1 DOPT_B200_PDL=0 timeout 200 python bench.py --timeline $OUT/${TAG}_timeline_serial.txt --no-cpu-baseline > /dev/null 2>> $OUT/${TAG}.err
head -12 $OUT/${TAG}_timeline_serial.txt
grep -A18 "idle time by" $OUT/${TAG}_timeline_serial.txt | cut -c1-3000
timeout 200 python bench.py --timeline $OUT/${TAG}_timeline.txt --no-cpu-baseline > /dev/null 2>> $OUT/${TAG}.err
grep -A16 "idle time by" $OUT/${TAG}_timeline.txt | head -17
timeout 900 python -m pytest tests/test_plan_gpu.py -q -x -k "28_10 or wrn_16 or interior" > $OUT/${TAG}_pytest_plan.log 2>&1
echo "pytest plan rc=$?" >> $OUT/${TAG}_pytest_plan.log
tail -4 $OUT/${TAG}_pytest_plan.log

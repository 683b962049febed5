#!/bin/bash
# round-2 GPU call Q: epilogue companions (residual sum in the convolution epilogue, backward batch-norm statistics in the
# feature-gradient epilogue): parity tests, bench with each switched off, serial timeline
set -u
OUT=gpurun_out
TAG=${1:-r02q}
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_plan_gpu.py -q -x -k "epilogue or 28_10 or wrn_16 or interior" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
tail -15 $OUT/${TAG}_pytest.log
run() {
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_${name}.json 2> $OUT/${TAG}_${name}.err
  python - <<PY
import json
f = "$OUT/${TAG}_bench_${name}.json"
try:
    d = json.load(open(f)); print("$name", round(d['value']), round(d['ms_per_step'], 3), round(d['e2e']['value']), d['roofline']['kernel_ms_per_step'], round(d['roofline']['frac_of_burst_peak'], 3), d['loss_first'], d['loss_last'], d['launches_per_step'], {k:(round(v['frac'],3), round(v['us_per_step'])) for k,v in d['roofline_classes'].items()})
    print(d['per_op_us_per_step'])
except Exception as e: print(f, "FAILED", e)
PY
  tail -3 $OUT/${TAG}_${name}.err
}
run both A=1
run noadd DOPT_B200_NO_EPI_ADD=1
run nobn DOPT_B200_NO_EPI_BNGRAD=1
run none DOPT_B200_NO_EPI_ADD=1 DOPT_B200_NO_EPI_BNGRAD=1
DOPT_B200_NO_SIDE_STREAM=1 DOPT_B200_PDL=0 timeout 200 python bench.py --timeline $OUT/${TAG}_timeline_serial.txt --no-cpu-baseline > /dev/null 2>> $OUT/${TAG}.err
head -30 $OUT/${TAG}_timeline_serial.txt
timeout 200 python bench.py --timeline $OUT/${TAG}_timeline.txt --no-cpu-baseline > /dev/null 2>> $OUT/${TAG}.err
head -8 $OUT/${TAG}_timeline.txt

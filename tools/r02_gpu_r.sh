#!/bin/bash
# round-2 GPU call R: filter-gradient scratches re-zeroed by the finish launch (no memsets), optimiser class = terminal launches
set -u
OUT=gpurun_out
TAG=${1:-r02r}
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_plan_gpu.py tests/test_cudnn_replay_gpu.py -q -x -k "28_10 or wrn_16 or interior or conv or passes or regular" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
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
run base A=1
run noarena DOPT_B200_NO_WG_ARENA=1
timeout 200 python bench.py --timeline $OUT/${TAG}_timeline.txt --no-cpu-baseline > /dev/null 2>> $OUT/${TAG}.err
head -8 $OUT/${TAG}_timeline.txt

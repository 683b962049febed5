#!/bin/bash
# round-2 GPU call K: side-stream filter gradients (A/B against the single-stream order on the same box), plan parity,
# ncu --set full captures of the convolution kernel (L160 fwd, L160 wgrad+dgrad, L640 dgrad+wgrad) and the flat batch norm
set -u
OUT=gpurun_out
TAG=${1:-r02k}
mkdir -p $OUT
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}.err
DOPT_B200_NO_SIDE_STREAM=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_noside.json 2>> $OUT/${TAG}.err
python - <<PY
import json
for f in ("$OUT/${TAG}_bench.json", "$OUT/${TAG}_bench_noside.json"):
    try:
        d = json.load(open(f)); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['loss_first'], d['loss_last'], d['launches_per_step'])
    except Exception as e: print(f, "FAILED", e)
PY
tail -5 $OUT/${TAG}.err
timeout 200 python bench.py --timeline $OUT/${TAG}_timeline.txt --no-cpu-baseline > /dev/null 2>> $OUT/${TAG}.err
head -14 $OUT/${TAG}_timeline.txt
timeout 900 python -m pytest tests/test_plan_gpu.py tests/test_ops_gpu.py -q -x -k "28_10 or wrn_16 or interior or passes or pointwise" > $OUT/${TAG}_pytest_plan.log 2>&1
echo "pytest plan rc=$?" >> $OUT/${TAG}_pytest_plan.log
tail -4 $OUT/${TAG}_pytest_plan.log
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:tc_kernel --launch-skip 178 --launch-count 3 -f -o $OUT/${TAG}_tc_fwd160 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> $OUT/${TAG}_ncu1.err
timeout 400 $NCU -k regex:tc_kernel --launch-skip 246 --launch-count 2 -f -o $OUT/${TAG}_tc_wgrad160 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> $OUT/${TAG}_ncu2.err
timeout 400 $NCU -k regex:tc_kernel --launch-skip 203 --launch-count 2 -f -o $OUT/${TAG}_tc_640 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> $OUT/${TAG}_ncu3.err
timeout 400 $NCU -k regex:flat_ --launch-skip 170 --launch-count 8 -f -o $OUT/${TAG}_flat python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> $OUT/${TAG}_ncu4.err
tail -2 $OUT/${TAG}_ncu1.err $OUT/${TAG}_ncu4.err
ls -la $OUT/*.ncu-rep

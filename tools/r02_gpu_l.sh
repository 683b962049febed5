#!/bin/bash
# round-2 GPU call L: halo pipeline of the 3x3 convolutions -- parity first, then bench / timeline, A/B against HALO=0
set -u
OUT=gpurun_out
TAG=${1:-r02l}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_cudnn_replay_gpu.py tests/test_ops_gpu.py -q -x -k "conv" > $OUT/${TAG}_pytest_conv.log 2>&1
echo "pytest conv rc=$?" >> $OUT/${TAG}_pytest_conv.log
tail -6 $OUT/${TAG}_pytest_conv.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}.err
DOPT_B200_HALO=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_halo1.json 2>> $OUT/${TAG}.err
DOPT_B200_HALO=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_halo0.json 2>> $OUT/${TAG}.err
DOPT_B200_CTAS_PER_SM=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_cta1.json 2>> $OUT/${TAG}.err
python - <<PY
import json
for f in ("$OUT/${TAG}_bench.json", "$OUT/${TAG}_bench_halo1.json", "$OUT/${TAG}_bench_halo0.json", "$OUT/${TAG}_bench_cta1.json"):
    try:
        d = json.load(open(f)); print(f, d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['frac_of_burst_peak'], d['loss_first'], d['loss_last'])
    except Exception as e: print(f, "FAILED", e)
PY
tail -5 $OUT/${TAG}.err
timeout 200 python bench.py --timeline $OUT/${TAG}_timeline.txt --no-cpu-baseline > /dev/null 2>> $OUT/${TAG}.err
head -12 $OUT/${TAG}_timeline.txt
tail -2 $OUT/${TAG}_timeline.txt | cut -c1-3000
timeout 900 python -m pytest tests/test_plan_gpu.py -q -x -k "28_10 or wrn_16 or interior or regularised" > $OUT/${TAG}_pytest_plan.log 2>&1
echo "pytest plan rc=$?" >> $OUT/${TAG}_pytest_plan.log
tail -6 $OUT/${TAG}_pytest_plan.log

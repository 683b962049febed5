#!/bin/bash
# round-2 GPU call M: programmatic dependent launch on the flat batch-norm kernels and the tcgen05 kernel -- full gpu suite, A/B bench
set -u
OUT=gpurun_out
TAG=${1:-r02m}
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
tail -6 $OUT/${TAG}_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}.err
DOPT_B200_PDL=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_nopdl.json 2>> $OUT/${TAG}.err
python - <<PY
import json
for f in ("$OUT/${TAG}_bench.json", "$OUT/${TAG}_bench_nopdl.json"):
    try:
        d = json.load(open(f)); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms_per_step'], d['roofline']['frac_of_burst_peak'], d['loss_first'], d['loss_last'])
    except Exception as e: print(f, "FAILED", e)
PY
tail -5 $OUT/${TAG}.err
timeout 200 python bench.py --timeline $OUT/${TAG}_timeline.txt --no-cpu-baseline > /dev/null 2>> $OUT/${TAG}.err
head -12 $OUT/${TAG}_timeline.txt

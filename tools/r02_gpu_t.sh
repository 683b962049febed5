#!/bin/bash
# sweep: CTAs per SM of the flat batch-norm / add kernels
set -u
OUT=gpurun_out
TAG=${1:-r02t}
mkdir -p $OUT
for v in 0 2 3 4 6 8; do
if [ $v = 0 ]; then E="A=1"; else E="DOPT_B200_FLAT_CTAS=$v"; fi
env $E timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_$v.json 2> $OUT/${TAG}_$v.err
python - <<PY
import json
f = "$OUT/${TAG}_bench_$v.json"
try:
    d = json.load(open(f)); print("ctas $v", round(d['value']), round(d['ms_per_step'], 3), {k:(round(v['frac'],3), round(v['us_per_step'])) for k,v in d['roofline_classes'].items()})
except Exception as e: print(f, "FAILED", e)
PY
done

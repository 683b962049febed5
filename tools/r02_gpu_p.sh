#!/bin/bash
# round-2 GPU call P: bench line with class-replay rooflines + compute-sanitizer
set -u
OUT=gpurun_out
TAG=${1:-r02p}
mkdir -p $OUT
timeout 400 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}.err
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench.json"))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['loss_first'], d['loss_last'])
print(d['roofline'])
print({k:(round(v['frac'],3), round(v['us_per_step'],1)) for k,v in d['roofline_classes'].items()})
print(d['cpu_baseline'], d['clocks'])
PY
tail -5 $OUT/${TAG}.err
bash tools/sanitize.sh $TAG

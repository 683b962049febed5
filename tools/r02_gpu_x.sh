#!/bin/bash
# check after the in-place write-back overlap fix: plan / optimiser tests, launches per step and step time unchanged
set -u
OUT=gpurun_out
TAG=${1:-r02x}
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_plan_gpu.py tests/test_optim_gpu.py -q -x > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
tail -4 $OUT/${TAG}_pytest.log
for i in 1 2 3 4; do
timeout 600 python -m pytest tests/test_plan_gpu.py -q -x -s -k "epilogue or interior" > $OUT/${TAG}_pytest_rep$i.log 2>&1
echo "rep $i rc=$?"; grep -E "epilogue:|passed|failed" $OUT/${TAG}_pytest_rep$i.log | cut -c1-220
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}.err
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench.json"))
print(round(d['value']), round(d['ms_per_step'], 3), round(d['e2e']['value']), d['launches_per_step'], d['loss_first'], d['loss_last'])
PY

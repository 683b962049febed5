#!/bin/bash
# test-time plans after relu + NHWC staging were absorbed into batchNormInference: parity in every mode, then the inference leg
set -u
OUT=gpurun_out
TAG=${1:-r02z6}
mkdir -p $OUT
timeout -k 5 200 python -m pytest tests/test_baseline_configs_gpu.py tests/test_plan_gpu.py -q -rf -k "test_time or inference" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
grep -E "^FAILED|^E  |passed|failed|rc=" $OUT/${TAG}_pytest.log | cut -c1-300 | head
for v in absorb noabsorb; do
if [ $v = noabsorb ]; then export DOPT_B200_NO_INFER_ABSORB=1; fi
timeout -k 5 120 python - > $OUT/${TAG}_infer_$v.json 2>$OUT/${TAG}_infer_$v.err <<PY
import json, torch
import bench
import dopt_b200 as db
from dopt_b200 import host as H
assert H.init(), H.init_error()
x, y, net, upd = bench.build_wrn(H, 128)
out = bench.adjacent_rows(torch, db, H, x, net, 128, bench.peaks()["hbm_gbs"])
print(json.dumps(out["inference"]))
PY
echo $v; cat $OUT/${TAG}_infer_$v.json | cut -c1-400; tail -2 $OUT/${TAG}_infer_$v.err
done

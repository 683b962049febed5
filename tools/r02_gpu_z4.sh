#!/bin/bash
# input pipeline after the 32-bit index arithmetic: bit-exactness tests, then the timing leg of bench.py on its own
set -u
OUT=gpurun_out
TAG=${1:-r02z4}
mkdir -p $OUT
timeout -k 5 200 python -m pytest tests/test_input_pipeline.py -q -rf > $OUT/${TAG}_pytest_input.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest_input.log
grep -E "^FAILED|^E  |passed|failed|rc=" $OUT/${TAG}_pytest_input.log | cut -c1-300 | head
timeout -k 5 120 python - > $OUT/${TAG}_input_timing.json 2>$OUT/${TAG}_input_timing.err <<PY
import json, torch
import bench
import dopt_b200 as db
class _N: pass
out = bench.adjacent_rows(torch, db, None, None, _N(), 128, bench.peaks()["hbm_gbs"])
print(json.dumps(out["input_pipeline"]))
PY
cat $OUT/${TAG}_input_timing.json | cut -c1-900; tail -2 $OUT/${TAG}_input_timing.err

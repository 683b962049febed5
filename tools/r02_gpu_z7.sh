#!/bin/bash
set -u
OUT=gpurun_out
TAG=${1:-r02z7}
mkdir -p $OUT
timeout -k 5 200 python -m pytest tests/test_baseline_configs_gpu.py -q -rf -k "mnist" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
grep -E "^FAILED|^E  |passed|failed|rc=" $OUT/${TAG}_pytest.log | cut -c1-400 | head -20

#!/bin/bash
set -u
OUT=gpurun_out
TAG=${1:-r02z2}
mkdir -p $OUT
timeout -k 5 200 python -m pytest tests/test_baseline_configs_gpu.py -q -rf -s -k vgg19 > $OUT/${TAG}_pytest_cfg.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest_cfg.log
grep -E "VGG19|convolution filters|^FAILED|^E  |passed|failed|rc=" $OUT/${TAG}_pytest_cfg.log | cut -c1-700 | head -40

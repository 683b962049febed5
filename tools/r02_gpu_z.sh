#!/bin/bash
# BASELINE configs other than the benchmarked one at full size (VGG19+BN N=100, SINS WRN-16-8 N=50 96x96, MNIST CNN): per-op parity
# against the replayed cuDNN calls, mode-invariance of a full-size training step, the test-time plan; bench line with the
# SURVEY 8(f) legs (input pipeline, inference)
set -u
OUT=gpurun_out
TAG=${1:-r02z}
mkdir -p $OUT
timeout -k 5 240 python -m pytest tests/test_cudnn_replay_gpu.py -q -rf -k "convolution_family or batchnorm_vs_cudnn or maxpool_and_grad" > $OUT/${TAG}_pytest_ops.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest_ops.log
grep -E "^FAILED|passed|failed|rc=" $OUT/${TAG}_pytest_ops.log | cut -c1-400 | head -30
timeout -k 5 400 python -m pytest tests/test_baseline_configs_gpu.py -q -rf -s > $OUT/${TAG}_pytest_cfg.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest_cfg.log
grep -E "first-step|AMSGrad losses|^FAILED|^E  |passed|failed|rc=" $OUT/${TAG}_pytest_cfg.log | cut -c1-600 | head -40
timeout -k 5 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/${TAG}_bench.json") if l.startswith("{")][-1])
    print(round(d['value']), round(d['ms_per_step'], 3), round(d['e2e']['value']), d['launches_per_step'], d['loss_first'], d['loss_last'])
    print(json.dumps(d.get('adjacent_rows'))[:1500])
except Exception as e:
    print("bench FAILED", e)
PY
tail -3 $OUT/${TAG}_bench.err | cut -c1-300

#!/bin/bash
# round-2 GPU call O: issue warps above the epilogue warps, L2 residency hints in the flat batch norm, delay kernel in the profile
set -u
OUT=gpurun_out
TAG=${1:-r02o}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_cudnn_replay_gpu.py tests/test_ops_gpu.py tests/test_plan_gpu.py -q -x -k "conv or batchnorm or 28_10 or wrn_16 or interior or matmul" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}.err
DOPT_B200_WG_PIX=128 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_wgpix128.json 2>> $OUT/${TAG}.err
python - <<PY
import json
for f in ("$OUT/${TAG}_bench.json", "$OUT/${TAG}_bench_wgpix128.json"):
    try:
        d = json.load(open(f)); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms_per_step'], d['roofline']['frac_of_burst_peak'], d['loss_first'], d['loss_last'], {k:(round(v['frac'],3), v['us_per_step']) for k,v in d['roofline_classes'].items()})
    except Exception as e: print(f, "FAILED", e)
PY
tail -5 $OUT/${TAG}.err
DOPT_B200_NO_SIDE_STREAM=1 DOPT_B200_PDL=0 timeout 200 python bench.py --timeline $OUT/${TAG}_timeline_serial.txt --no-cpu-baseline > /dev/null 2>> $OUT/${TAG}.err
head -14 $OUT/${TAG}_timeline_serial.txt
DOPT_B200_WG_PIX=128 DOPT_B200_NO_SIDE_STREAM=1 DOPT_B200_PDL=0 timeout 200 python bench.py --timeline $OUT/${TAG}_timeline_serial_wgpix128.txt --no-cpu-baseline > /dev/null 2>> $OUT/${TAG}.err
python - <<PY
import re
for f in ("$OUT/${TAG}_timeline_serial.txt", "$OUT/${TAG}_timeline_serial_wgpix128.txt"):
    t = open(f).read().split('in order:\n')[1]
    items = re.findall(r'<([^>]*)>:(\d+)', t)
    print(' '.join(('W' if k.startswith('2') else ('S' if k.split(', ')[4] == 'true' else 'C')) + ('h' if k.split(', ')[5] == 'true' else '') + ':' + d for k, d in items))
    print(sum(int(d) for k, d in items))
PY

#!/bin/bash
# round-2 final evidence on one B200: smoke(), the whole gpu test suite, the parity figures the tests print, the bench line,
# ncu launch list / tensor-core DRAM traffic / CUPTI timelines (tools/capture_profiles.sh) and ncu --set full captures of the
# convolution kernels of the 160-channel layers (forward halo pipeline with and without the statistics epilogue, filter gradient)
set -u
OUT=gpurun_out
TAG=${1:-r02final}
mkdir -p $OUT
timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1
echo "smoke rc=$?" >> $OUT/${TAG}_smoke.log
tail -2 $OUT/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q -rs > $OUT/${TAG}_pytest.log 2>&1
echo "rc=$?" >> $OUT/${TAG}_pytest.log
tail -4 $OUT/${TAG}_pytest.log
timeout 900 python -m pytest tests/test_plan_gpu.py -q -s -k "28_10 or wrn_16_4 or epilogue" > $OUT/${TAG}_parity.log 2>&1
grep -E "losses|update errors|epilogue:|passed|failed" $OUT/${TAG}_parity.log | cut -c1-400
bash tools/capture_profiles.sh $TAG
DOPT_B200_NO_SIDE_STREAM=1 DOPT_B200_PDL=0 timeout 200 python bench.py --timeline $OUT/${TAG}_timeline_serial.txt --no-cpu-baseline > /dev/null 2>> $OUT/${TAG}.err
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench.json"))
print(round(d['value']), round(d['ms_per_step'], 3), round(d['e2e']['value']), d['roofline']['kernel_ms_per_step'], round(d['roofline']['frac_of_burst_peak'], 3), round(d['roofline']['frac'], 3), d['loss_first'], d['loss_last'], d['launches_per_step'], {k:(round(v['frac'],3), round(v['us_per_step'])) for k,v in d['roofline_classes'].items()})
print(d['cpu_baseline'], d['clocks'])
PY
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:tc_kernel --launch-skip 178 --launch-count 3 -f -o $OUT/${TAG}_tc_fwd160 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> $OUT/${TAG}_ncu1.err
timeout 400 $NCU -k regex:tc_kernel --launch-skip 246 --launch-count 3 -f -o $OUT/${TAG}_tc_bwd160 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> $OUT/${TAG}_ncu2.err
tail -2 $OUT/${TAG}_ncu1.err $OUT/${TAG}_ncu2.err
ls -la $OUT/${TAG}*.ncu-rep

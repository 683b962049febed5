#!/bin/bash
# One-call capture of the evidence profiles/ holds for a round (run on the GPU box from the repo root):
#
#   gpurun --timeout 600 -- 'bash tools/capture_profiles.sh r02a'
#
# Writes into gpurun_out/ (copy what should be judged into profiles/):
#   <tag>_bench.json       bench line of the build (never taken under a profiler)
#   <tag>_launches.csv     ncu launch list (gpu__time_duration.sum) of about one graph-replayed training step
#   <tag>_launches.md      per-kernel shares of that list (tools/ncu_summary.py)
#   <tag>_tc_dram.csv      dram__bytes_read/write of the tensor-core launches of one step (bench.py's roofline.traffic)
#   <tag>_timeline.txt     CUPTI per-kernel summary of three graph-replayed steps (device time, launch count, idle gaps)
# The launch window: warm-up and graph capture issue ~1100 launches before the first replayed step; LAUNCHES is the number
# of launches per step the bench line reports (launches_per_step), read from the fresh bench run.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
timeout 120 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
LAUNCHES=$(python -c "import json;print(int(json.load(open('$OUT/${TAG}_bench.json'))['launches_per_step']))" 2>/dev/null || echo 300)
SKIP=${SKIP:-1100}
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip $SKIP --launch-count $((LAUNCHES + 30)) \
    --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python tools/ncu_summary.py $OUT/${TAG}_launches.csv "$TAG: launches of about one graph-replayed WRN-28-10 step" > $OUT/${TAG}_launches.md 2>/dev/null
timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:tc_kernel --launch-skip 300 \
    --launch-count 88 --csv --log-file $OUT/${TAG}_tc_dram.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 100 python bench.py --timeline $OUT/${TAG}_timeline.txt --no-cpu-baseline > /dev/null 2>&1
ls -la $OUT | grep ${TAG}

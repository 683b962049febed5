#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 200 python bench.py --timeline $OUT/r02b_timeline.txt --no-cpu-baseline > /dev/null 2> $OUT/r02b.err
head -45 $OUT/r02b_timeline.txt
for c in 1 2 3 6 8; do
  DOPT_B200_FLAT_CTAS=$c timeout 200 python bench.py --timeline $OUT/r02b_timeline_c$c.txt --no-cpu-baseline > /dev/null 2>> $OUT/r02b.err
  echo "== FLAT_CTAS=$c"; grep -E "busy|flat_" $OUT/r02b_timeline_c$c.txt
done

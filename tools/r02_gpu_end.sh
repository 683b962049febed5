#!/bin/bash
# evidence of the round's last build, most important first: gpu suite (as the driver runs it), smoke(), the bench line, then the
# ncu launch list of one graph-replayed step and the CUPTI timeline
set -u
OUT=gpurun_out
TAG=${1:-r02end}
mkdir -p $OUT
timeout -k 5 400 python -m pytest tests -m gpu -x -q -rs > $OUT/${TAG}_pytest.log 2>&1
echo "rc=$?" >> $OUT/${TAG}_pytest.log
tail -4 $OUT/${TAG}_pytest.log | cut -c1-200
timeout -k 5 100 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
echo "smoke rc=$?" >> $OUT/${TAG}_smoke.log
tail -2 $OUT/${TAG}_smoke.log | cut -c1-200
timeout -k 5 200 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/${TAG}_bench.json") if l.startswith("{")][-1])
    print(round(d['value']), round(d['ms_per_step'], 3), round(d['e2e']['value']), d['launches_per_step'], d['loss_first'], d['loss_last'],
          round(d['roofline']['frac_of_burst_peak'], 3), d['clocks'], d['cpu_baseline'] and round(d['cpu_baseline']['value'], 2))
    print(json.dumps(d.get('adjacent_rows'))[:900])
except Exception as e:
    print("bench FAILED", e)
PY
LAUNCHES=$(python -c "import json;print(int(json.loads([l for l in open('$OUT/${TAG}_bench.json') if l.startswith('{')][-1])['launches_per_step']))" 2>/dev/null || echo 227)
timeout -k 5 200 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip ${SKIP:-1100} --launch-count $((LAUNCHES + 30)) \
    --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-adjacent > /dev/null 2>&1
python tools/ncu_summary.py $OUT/${TAG}_launches.csv "$TAG: launches of about one graph-replayed WRN-28-10 step" > $OUT/${TAG}_launches.md 2>/dev/null
head -12 $OUT/${TAG}_launches.md | cut -c1-120
timeout -k 5 100 python bench.py --timeline $OUT/${TAG}_timeline.txt --no-cpu-baseline --no-adjacent > /dev/null 2>&1
head -3 $OUT/${TAG}_timeline.txt | cut -c1-160

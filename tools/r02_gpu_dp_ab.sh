#!/bin/bash
# 2-GPU A/B: which of the round-2 changes costs data-parallel efficiency
set -u
OUT=gpurun_out
TAG=${1:-r02dpab}
mkdir -p $OUT
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus 2 --steps 20 --warmup 3 > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/${TAG}_${name}.json") if l.startswith("{")][-1])
    print("$name", round(d["value"]), round(d["ms_per_step"], 3), d["per_op_us_per_step"].get("allreduceBucket"))
except Exception as e:
    print("$name", "FAILED", e)
PY
}
run base A=1
run noside DOPT_B200_NO_SIDE_STREAM=1
run nopdl DOPT_B200_PDL=0
run noside_nopdl DOPT_B200_NO_SIDE_STREAM=1 DOPT_B200_PDL=0
run nohalo DOPT_B200_HALO=0 DOPT_B200_WG_HALO=0
env timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29777 bench.py --gpus 2 --timeline $OUT/${TAG}_timeline.txt > /dev/null 2> $OUT/${TAG}_timeline.err
head -24 $OUT/${TAG}_timeline.txt
grep -A12 "idle time by" $OUT/${TAG}_timeline.txt

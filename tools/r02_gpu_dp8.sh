#!/bin/bash
# 8-GPU: gate on/off and channel count
set -u
OUT=gpurun_out
NG=${1:-8}
TAG=${2:-r02dp8}
mkdir -p $OUT
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $NG --steps 20 --warmup 3 > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/${TAG}_${name}.json") if l.startswith("{")][-1])
    print("$name", round(d["value"]), round(d["ms_per_step"], 3), d["per_op_us_per_step"].get("allreduceBucket"), d["per_op_us_per_step"].get("fusedRegion"), d["launches_per_step"])
except Exception as e:
    print("$name", "FAILED", e)
PY
  grep -i "error\|Traceback" $OUT/${TAG}_${name}.err | head -3
}
run base NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING
run gate_off DOPT_B200_GATE_SMS=0
run ch8_nogate DOPT_B200_COMM_CHANNELS=8 DOPT_B200_GATE_SMS=0
run ch32_nogate DOPT_B200_COMM_CHANNELS=32 DOPT_B200_GATE_SMS=0
run ch32 DOPT_B200_COMM_CHANNELS=32
grep -i "nvls\|algo\|proto\|channels" $OUT/${TAG}_base.err | head -12
env timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29777 bench.py --gpus $NG --timeline $OUT/${TAG}_timeline.txt > /dev/null 2> $OUT/${TAG}_timeline.err
head -8 $OUT/${TAG}_timeline.txt

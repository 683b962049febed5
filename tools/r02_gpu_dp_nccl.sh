#!/bin/bash
# N-GPU sweep of the NCCL knobs behind the gradient-bucket all-reduce (channels = reserved SMs, protocol)
set -u
OUT=gpurun_out
NG=${1:-2}
TAG=${2:-r02nccl}
mkdir -p $OUT
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $NG --steps 20 --warmup 3 > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/${TAG}_${name}.json") if l.startswith("{")][-1])
    print("$name", round(d["value"]), round(d["ms_per_step"], 3), d["per_op_us_per_step"].get("allreduceBucket"))
except Exception as e:
    print("$name", "FAILED", e)
PY
}
run ch4 A=1
run ch0 DOPT_B200_COMM_CHANNELS=0
run ch8 DOPT_B200_COMM_CHANNELS=8
run ch16 DOPT_B200_COMM_CHANNELS=16
run ch4_simple NCCL_PROTO=Simple
run ch8_simple DOPT_B200_COMM_CHANNELS=8 NCCL_PROTO=Simple
run ch16_simple DOPT_B200_COMM_CHANNELS=16 NCCL_PROTO=Simple
run ch0_simple DOPT_B200_COMM_CHANNELS=0 NCCL_PROTO=Simple

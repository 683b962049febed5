#!/bin/bash
set -u
NG=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
run() {
  local tag=$1; shift
  env PROBE_TAG=$tag "$@" timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) tools/nccl_probe.py 2> $OUT/probe_${NG}_${tag}.err | grep world
  grep -i "error\|Traceback" $OUT/probe_${NG}_${tag}.err | head -2
}
if [ "$NG" = "8" ]; then
run ch16 A=1
run ch16_nvls NCCL_ALGO=NVLS
run ch16_simple NCCL_PROTO=Simple
run ch8_nvls DOPT_B200_COMM_CHANNELS=8 NCCL_ALGO=NVLS
run ch32 DOPT_B200_COMM_CHANNELS=32
run ch32_nvls DOPT_B200_COMM_CHANNELS=32 NCCL_ALGO=NVLS
run free DOPT_B200_COMM_CHANNELS=0
exit 0
fi
run ch16 A=1
run ch16_simple NCCL_PROTO=Simple
run ch16_ll128 NCCL_PROTO=LL128
run ch16_nvls NCCL_ALGO=NVLS
run ch16_tree NCCL_ALGO=Tree
run ch32 DOPT_B200_COMM_CHANNELS=32
run ch32_simple DOPT_B200_COMM_CHANNELS=32 NCCL_PROTO=Simple
run ch8_simple DOPT_B200_COMM_CHANNELS=8 NCCL_PROTO=Simple
run free DOPT_B200_COMM_CHANNELS=0
run free_nvls DOPT_B200_COMM_CHANNELS=0 NCCL_ALGO=NVLS

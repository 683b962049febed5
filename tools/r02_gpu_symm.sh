#!/bin/bash
set -u
NG=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29611 tools/symm_probe.py > $OUT/symm_probe_$NG.log 2>&1
echo "rc=$?"; grep -v "OMP_NUM\|\*\*\*\*" $OUT/symm_probe_$NG.log | tail -15
PROBE_ALL_VISIBLE=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29612 tools/symm_probe.py > $OUT/symm_probe_${NG}_allvis.log 2>&1
echo "rc=$?"; grep -v "OMP_NUM\|\*\*\*\*" $OUT/symm_probe_${NG}_allvis.log | tail -15

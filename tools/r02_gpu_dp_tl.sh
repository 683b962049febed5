#!/bin/bash
# per-stream kernel timeline of one data-parallel step and of one single-GPU step
set -u
OUT=gpurun_out
NG=${1:-2}
TAG=${2:-r02dptl}
mkdir -p $OUT
timeout 300 python bench.py --timeline $OUT/${TAG}_n1.txt --no-cpu-baseline > /dev/null 2> $OUT/${TAG}_n1.err
env timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29777 bench.py --gpus $NG --timeline $OUT/${TAG}_n${NG}.txt > /dev/null 2> $OUT/${TAG}_n${NG}.err
wc -l $OUT/${TAG}_n1.txt.raw $OUT/${TAG}_n${NG}.txt.raw
tail -3 $OUT/${TAG}_n1.err $OUT/${TAG}_n${NG}.err

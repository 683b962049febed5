#!/bin/bash
# multi-GPU call: (N=2) the NCCL parity check of tools/dp_check.py with its log, then bench.py at every rank count up to N
# usage: gpurun --gpus N -- 'bash tools/r02_gpu_dp.sh N TAG'
set -u
N=${1:-2}
TAG=${2:-r02dp}
OUT=gpurun_out
mkdir -p $OUT
if [ "$N" -ge 2 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py > $OUT/${TAG}_dp_check.log 2>&1
  echo "dp_check rc=$?" >> $OUT/${TAG}_dp_check.log
  grep -E "PASS|FAIL|rc=" $OUT/${TAG}_dp_check.log | tail -8
fi
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_n1.err
for n in 2 4 8; do
  if [ "$n" -le "$N" ]; then
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520 + n)) bench.py --gpus $n --steps 20 --warmup 3 > $OUT/${TAG}_bench_n$n.json 2> $OUT/${TAG}_n$n.err
  fi
done
python - <<PY
import json, glob
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.loads([l for l in open("$OUT/${TAG}_bench_n%d.json" % n) if l.startswith("{")][-1])
    except Exception as e:
        continue
    if n == 1: base = d["value"]
    print(n, round(d["value"]), round(d["ms_per_step"], 3), "eff", round(d["value"] / (n * base), 4) if base else None, d.get("clocks"))
PY

#!/bin/bash
# N-GPU: data-parallel correctness (NCCL parity test), then the bench with the round-2 exchange changes switched off one by one
set -u
OUT=gpurun_out
NG=${1:-2}
TAG=${2:-r02dp3}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_dp_nccl_gpu.py -q -x > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
tail -4 $OUT/${TAG}_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_n1.json 2> $OUT/${TAG}_n1.err
python - <<PY
import json
d = json.loads([l for l in open("$OUT/${TAG}_n1.json") if l.startswith("{")][-1])
print("n1", round(d["value"]), round(d["ms_per_step"], 3))
PY
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $NG --steps 20 --warmup 3 > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/${TAG}_${name}.json") if l.startswith("{")][-1])
    print("$name", round(d["value"]), round(d["ms_per_step"], 3), d["per_op_us_per_step"].get("allreduceBucket"), d["per_op_us_per_step"].get("fusedRegion"), d["launches_per_step"])
except Exception as e:
    print("$name", "FAILED", e)
PY
  tail -2 $OUT/${TAG}_${name}.err
}
run base A=1
run nosidefinish DOPT_B200_NO_SIDE_FINISH=1
run gate_off DOPT_B200_GATE_SMS=0
run noside DOPT_B200_NO_SIDE_STREAM=1
run ch8 DOPT_B200_COMM_CHANNELS=8
run ch32 DOPT_B200_COMM_CHANNELS=32
env timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29777 bench.py --gpus $NG --timeline $OUT/${TAG}_timeline.txt > /dev/null 2> $OUT/${TAG}_timeline.err
head -12 $OUT/${TAG}_timeline.txt

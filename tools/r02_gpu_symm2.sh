#!/bin/bash
# N-GPU: the library's own all-reduce over NVSwitch multicast against ncclAllReduce -- parity test, then the bench both ways
set -u
OUT=gpurun_out
NG=${1:-2}
TAG=${2:-r02symm}
mkdir -p $OUT
if [ "$NG" = "2" ]; then
timeout 500 python -m pytest tests/test_dp_nccl_gpu.py -q -x > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
tail -12 $OUT/${TAG}_pytest.log | cut -c1-300
cat $OUT/dp_check_multicast.log 2>/dev/null | grep -v "OMP_NUM\|\*\*\*\*" | tail -5 | cut -c1-300
fi
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $NG --steps 20 --warmup 3 > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/${TAG}_${name}.json") if l.startswith("{")][-1])
    print("$name", round(d["value"]), round(d["ms_per_step"], 3), d["per_op_us_per_step"].get("allreduceBucket"), d["launches_per_step"], d["loss_first"], d["loss_last"], (d["config"].get("exchange") or "")[:40])
except Exception as e:
    print("$name", "FAILED", e)
PY
  grep -i "error\|Traceback\|symm" $OUT/${TAG}_${name}.err | head -4
}
if [ "$NG" = "8" ]; then
run multicast A=1
run nccl DOPT_B200_SYMM=0
run multicast_c32 DOPT_B200_NVLS_CTAS=32
exit 0
fi
run multicast A=1
run nccl DOPT_B200_SYMM=0
run multicast_c8 DOPT_B200_NVLS_CTAS=8
run multicast_c32 DOPT_B200_NVLS_CTAS=32
run multicast2 A=1

#!/bin/bash
# final checks after the multicast all-reduce went in.  NG=1: smoke, whole gpu suite, bench.  NG=2: NCCL inside the all-visible
# process layout (symmetric pool attached but not used), N=2 bench.  NG=4: the N=4 point of the scaling table.
set -u
OUT=gpurun_out
NG=${1:-1}
TAG=${2:-r02w}
mkdir -p $OUT
if [ "$NG" = "1" ]; then
  timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/${TAG}_smoke.log; tail -2 $OUT/${TAG}_smoke.log
  timeout 1500 python -m pytest tests -m gpu -q -rs > $OUT/${TAG}_pytest.log 2>&1; echo "rc=$?" >> $OUT/${TAG}_pytest.log; tail -4 $OUT/${TAG}_pytest.log
  timeout 300 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_n1.err
  python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench_n1.json"))
print("n1", round(d['value']), round(d['ms_per_step'], 3), round(d['e2e']['value']), round(d['roofline']['frac_of_burst_peak'], 3), round(d['roofline']['frac'], 3), d['launches_per_step'], {k:round(v['frac'],3) for k,v in d['roofline_classes'].items()})
PY
  timeout 200 python bench.py --impl reference --steps 1 --warmup 0 | cut -c1-300
  exit 0
fi
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $NG --steps 20 --warmup 3 > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/${TAG}_${name}.json") if l.startswith("{")][-1])
    print("$name", round(d["value"]), round(d["ms_per_step"], 3), round(d["e2e"]["value"]), d["per_op_us_per_step"].get("allreduceBucket"), d["launches_per_step"], (d["config"].get("exchange") or "")[:40])
except Exception as e:
    print("$name", "FAILED", e)
PY
  grep -i "error\|Traceback\|symm" $OUT/${TAG}_${name}.err | head -4
}
run n${NG}_multicast A=1
if [ "$NG" = "2" ]; then
  run n2_nonvls DOPT_B200_NO_NVLS=1
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 2>/dev/null | cut -c1-300
fi

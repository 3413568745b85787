#!/bin/bash
# round-2 GPU call D2 (N GPUs, default 2): fused layout switches at the REAL level shapes (tools/peer_fused_probe.py: aligned levels 0-2 first,
# the straddling level 3 last and alone), then bench.py fused ("aligned" default) with pure frame sharding.
N=${1:-2}
EXTRA_BENCH=${2:---no-cfg-split}
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
exec > >(tee $O/d2_stdout_n$N.txt) 2>&1
T0=$(date +%s)
stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
stamp start
VC_PROBE_LEVELS=0,1,2 timeout 120 $TR --master-port 29541 tools/peer_fused_probe.py > $O/d2_probe_aligned_n$N.log 2>&1; RCA=$?; echo "probe levels 0,1,2 rc=$RCA"
grep -E "level|PROBE|Error|error" $O/d2_probe_aligned_n$N.log | grep -v "^W0" | tail -34
stamp probe_aligned
if [ "$N" != "2" ]; then
  VC_PEER_COMM=1 timeout 150 $TR --master-port 29544 tools/parallel_check.py > $O/d2_check_n$N.log 2>&1; echo "parallel_check N=$N rc=$?"
  grep -E "world|fused switches|PARALLEL_CHECK_OK|Error|error|Traceback" $O/d2_check_n$N.log | tail -16
  stamp parallel_check
fi
run() { # name env...
  local name=$1; shift
  env "$@" timeout 200 $TR --master-port 29542 bench.py --gpus $N --steps 6 --warmup 3 $EXTRA_BENCH > $O/d2_bench_n${N}_$name.json 2> $O/d2_bench_n${N}_$name.err
  echo "bench N=$N $name rc=$? $(cut -c1-150 $O/d2_bench_n${N}_$name.json)"; grep -E "Error|error" $O/d2_bench_n${N}_$name.err | tail -3
  python - <<PY
import json
try:
    d = json.loads(open("$O/d2_bench_n${N}_$name.json").read().strip().splitlines()[-1])
    print("   comm:", json.dumps(d.get("comm"))[:700]); print("   shard err:", d.get("sharded_vs_single_max_err"), "launches", d.get("gpu_launches"), "e2e", d.get("e2e", {}).get("value"))
except Exception as e:
    print("   (no json)", e)
PY
}
if [ $RCA = 0 ]; then run fused_aligned VC_PEER_FUSED=aligned; stamp bench_fused; fi
VC_PROBE_LEVELS=3 timeout 90 $TR --master-port 29543 tools/peer_fused_probe.py > $O/d2_probe_l3_n$N.log 2>&1; echo "probe level 3 (straddling patches) rc=$?"
grep -E "level|PROBE|rror" $O/d2_probe_l3_n$N.log | grep -v "^W0" | tail -14
stamp probe_l3

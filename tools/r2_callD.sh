#!/bin/bash
# round-2 GPU call D (N GPUs, default 2): layout switches fused into the producing GEMM's epilogue (VC_PEER_FUSED): parallel_check (unit checks
# fused vs separate exchange, sharded forward vs single GPU / oracle, graph replay), then bench.py with pure frame sharding (--no-cfg-split at
# N = 2 exercises the frame exchange) fused vs separate.
N=${1:-2}
EXTRA_BENCH=${2:---no-cfg-split}
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
exec > >(tee $O/d_stdout_n$N.txt) 2>&1
T0=$(date +%s)
stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi -L | head -8
stamp start
VC_PEER_COMM=1 VC_PEER_FUSED=1 timeout 150 $TR --master-port 29541 tools/parallel_check.py > $O/d_check_fused_n$N.log 2>&1; RC=$?; echo "parallel_check N=$N fused rc=$RC"
grep -E "world|peer exchange|fused switches|PARALLEL_CHECK_OK|Error|error|Traceback|failed" $O/d_check_fused_n$N.log | tail -24
stamp check_fused
run() { # name env...
  local name=$1; shift
  env "$@" timeout 200 $TR --master-port 29542 bench.py --gpus $N --steps 6 --warmup 3 $EXTRA_BENCH > $O/d_bench_n${N}_$name.json 2> $O/d_bench_n${N}_$name.err
  echo "bench N=$N $name rc=$? $(cut -c1-150 $O/d_bench_n${N}_$name.json)"; grep -E "Error|error" $O/d_bench_n${N}_$name.err | tail -3
  python - <<PY
import json
try:
    d = json.loads(open("$O/d_bench_n${N}_$name.json").read().strip().splitlines()[-1])
    print("   comm:", json.dumps(d.get("comm"))[:700]); print("   shard err:", d.get("sharded_vs_single_max_err"), "launches", d.get("gpu_launches"), "e2e", d.get("e2e", {}).get("value"))
except Exception as e:
    print("   (no json)", e)
PY
}
if [ $RC = 0 ]; then run fused VC_PEER_FUSED=1; stamp bench_fused; fi
run separate VC_PEER_FUSED=0
stamp bench_separate

#!/bin/bash
# round-2 GPU call C (N GPUs, default 4): frame-sharded U-Net over the NVLink peer-memory kernels vs NCCL vs single GPU (tools/parallel_check.py,
# P = N ranks in one frame group + the 2 x N/2 CFG-split step), then bench.py at N GPUs with both exchange implementations.
N=${1:-4}
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
exec > >(tee $O/c_stdout_n$N.txt) 2>&1
T0=$(date +%s)
stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi -L; nvidia-smi topo -m | head -12
stamp start
VC_DEBUG_SYNC=1 VC_PEER_COMM=1 timeout 240 $TR --master-port 29541 tools/parallel_check.py > $O/c_check_peer1_n$N.log 2>&1; echo "parallel_check N=$N peer=1 rc=$?"
grep -E "world|peer exchange|PARALLEL_CHECK_OK|Error|error|Traceback" $O/c_check_peer1_n$N.log | tail -14
stamp check_peer
run() { # name env...
  local name=$1; shift
  env "$@" timeout 300 $TR --master-port 29542 bench.py --gpus $N --steps 6 --warmup 3 > $O/c_bench_n${N}_$name.json 2> $O/c_bench_n${N}_$name.err
  echo "bench N=$N $name rc=$? $(cut -c1-150 $O/c_bench_n${N}_$name.json)"; tail -2 $O/c_bench_n${N}_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/c_bench_n${N}_$name.json").read().strip().splitlines()[-1])
    print("   comm:", json.dumps(d.get("comm"))[:600]); print("   shard err:", d.get("sharded_vs_single_max_err"), "launches", d.get("gpu_launches"), "e2e", d.get("e2e", {}).get("value"))
except Exception as e:
    print("   (no json)", e)
PY
}
run peer_graph VC_PEER_COMM=1
stamp bench_peer
run nccl_graph VC_PEER_COMM=0
stamp bench_nccl
VC_PEER_COMM=0 timeout 200 $TR --master-port 29543 tools/parallel_check.py > $O/c_check_peer0_n$N.log 2>&1; echo "parallel_check N=$N peer=0 rc=$?"
grep -E "world|PARALLEL_CHECK_OK|Error|error|Traceback" $O/c_check_peer0_n$N.log | tail -8
stamp check_nccl

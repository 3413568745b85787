#!/bin/bash
# multi-GPU evidence run: gpurun --gpus N -- 'bash tools/r2_multi.sh N'
# (1) tools/parallel_check.py with the peer-memory kernels and with NCCL, (2) bench.py at N GPUs: peer / NCCL, graph / eager
N=${1:-2}
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > $O/m${N}_topo.txt 2>&1
for peer in 1 0; do
  VC_DEBUG_SYNC=1 VC_PEER_COMM=$peer timeout 600 $TR --master-port 29541 tools/parallel_check.py > $O/m${N}_check_peer$peer.log 2>&1
  echo "parallel_check N=$N peer=$peer rc=$?"; grep -E "world|peer exchange|rank [0-9]\]|PARALLEL_CHECK_OK|Error|error" $O/m${N}_check_peer$peer.log | tail -12
done
run() { # name, env..., args
  local name=$1; shift
  env "$@" timeout 600 $TR --master-port 29542 bench.py --gpus $N --steps 6 --warmup 3 $EXTRA > $O/m${N}_bench_$name.json 2> $O/m${N}_bench_$name.err
  echo "bench N=$N $name rc=$? $(cut -c1-150 $O/m${N}_bench_$name.json)"; tail -2 $O/m${N}_bench_$name.err
}
EXTRA="" run peer_graph VC_PEER_COMM=1
EXTRA="--no-graph" run peer_eager VC_PEER_COMM=1
EXTRA="" run nccl_graph VC_PEER_COMM=0
EXTRA="--no-graph" run nccl_eager VC_PEER_COMM=0

#!/bin/bash
# round-2 GPU call D3 (2 GPUs): fused layout switches after the padding-tile fix (probe at the real shapes with poisoned receive buffers,
# bench with pure frame sharding fused vs the separate exchange measured in call D), then -- the two GPUs independently -- the 1-GPU suite on
# GPU 1 and the full default bench line on GPU 0.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
exec > >(tee $O/d3_stdout.txt) 2>&1
T0=$(date +%s)
stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
stamp start
VC_PROBE_LEVELS=0,1,2 timeout 120 $TR --master-port 29541 tools/peer_fused_probe.py > $O/d3_probe_aligned.log 2>&1; RCA=$?; echo "probe levels 0,1,2 rc=$RCA"
grep -E "level|PROBE" $O/d3_probe_aligned.log | grep -v "^W0" | grep -E "False|PROBE|EXCEPTION" | tail -12
grep -c "True" $O/d3_probe_aligned.log
stamp probe_aligned
if [ $RCA = 0 ]; then
  VC_PEER_FUSED=aligned timeout 200 $TR --master-port 29542 bench.py --gpus 2 --steps 6 --warmup 3 --no-cfg-split > $O/d3_bench_n2_fused.json 2> $O/d3_bench_n2_fused.err
  echo "bench N=2 fused rc=$? $(cut -c1-150 $O/d3_bench_n2_fused.json)"
  python - <<PY
import json
try:
    d = json.loads(open("$O/d3_bench_n2_fused.json").read().strip().splitlines()[-1])
    print("   comm:", json.dumps(d.get("comm"))[:700]); print("   shard err:", d.get("sharded_vs_single_max_err"), "launches", d.get("gpu_launches"), "e2e", d.get("e2e", {}).get("value"))
except Exception as e:
    print("   (no json)", e)
PY
  stamp bench_fused
fi
# NOTE: a bare `wait` also waits for the `tee` of the exec redirection above and never returns (this cost the rest of the round's GPU budget):
# wait for the job's PID only
( CUDA_VISIBLE_DEVICES=1 timeout 400 python -m pytest tests -m gpu -q -rf --deselect tests/test_multigpu_gpu.py > $O/d3_pytest.log 2>&1; echo "pytest rc=$?" >> $O/d3_pytest.log ) &
PYTEST_PID=$!
CUDA_VISIBLE_DEVICES=0 timeout 400 python bench.py --steps 6 --warmup 3 > $O/d3_bench_n1_full.json 2> $O/d3_bench_n1_full.err; echo "bench N=1 full rc=$? $(cut -c1-160 $O/d3_bench_n1_full.json)"
wait $PYTEST_PID
tail -6 $O/d3_pytest.log
python - <<PY
import json
try:
    d = json.loads(open("$O/d3_bench_n1_full.json").read().strip().splitlines()[-1])
    for k in ("value", "ms_per_step", "e2e", "gpu_launches", "clocks", "parity", "roofline_groupnorm", "roofline_attention", "gpu_eager_baseline", "vae_decode", "cpu_baseline_config1"):
        print("  ", k, json.dumps(d.get(k))[:400])
    r = d.get("roofline", {}); print("   roofline", r.get("achieved"), r.get("frac"), r.get("ms"), r.get("traffic"))
except Exception as e:
    print("   (no json)", e)
PY
stamp n1_jobs

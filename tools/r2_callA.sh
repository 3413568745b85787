#!/bin/bash
# round-2 GPU call A (1 GPU): GPU suite incl. the GroupNorm-from-producer tests, micro A/Bs (GN partial sums, ping-pong attention), bench A/Bs,
# ncu launch list + ncu --set full of the dominant kernels, config 2 / 5 bench lines, 50-step parity.  Outputs: gpurun_out/a_*
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
T0=$(date +%s)
stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*" | tee -a $O/a_timeline.txt; }
nvidia-smi -L > $O/a_smi.txt 2>&1
stamp start
timeout 1200 python -m pytest tests -m gpu -q -rf --deselect tests/test_multigpu_gpu.py > $O/a_pytest.log 2>&1; echo "pytest rc=$?" >> $O/a_pytest.log
tail -15 $O/a_pytest.log
stamp pytest
timeout 300 python tools/gn_parts_micro.py > $O/a_gn_parts_micro.txt 2>&1; cat $O/a_gn_parts_micro.txt | tail -30
stamp gn_micro
timeout 300 python tools/ab_micro.py > $O/a_ab_micro.txt 2>&1
VC_ATTN_PP=1 timeout 300 python tools/ab_micro.py 2>&1 | grep -E "attn|rror" >> $O/a_ab_micro.txt
VC_ATTN_PP=1 timeout 200 python tools/attn_check.py > $O/a_attn_check_pp.txt 2>&1; tail -3 $O/a_attn_check_pp.txt
grep -E "attn" $O/a_ab_micro.txt
stamp ab_micro
bench() { # name env... 
  local name=$1; shift
  env "$@" timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-vae > $O/a_bench_$name.json 2> $O/a_bench_$name.err
  echo "bench $name rc=$? $(cut -c1-140 $O/a_bench_$name.json)"
}
bench gn1 VC_GN_FROM_PRODUCER=1
bench gn0 VC_GN_FROM_PRODUCER=0
bench gn2 VC_GN_FROM_PRODUCER=2
bench gn1_pp VC_GN_FROM_PRODUCER=1 VC_ATTN_PP=1
stamp bench_ab
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_tap|flash_attn|gn_|temporal_attn|ln_|small_linear|im2col|upsample|ddim_|nchw|nhwc|cast_|timestep|peer_" \
  --csv --log-file $O/a_launches_bench.csv python bench.py --steps 1 --warmup 2 --no-graph --no-cpu-baseline --no-gpu-baseline --no-vae > $O/a_launches_bench.log 2>&1
python tools/launch_summary.py $O/a_launches_bench.csv > $O/a_launches_bench.summary.txt 2>&1; head -25 $O/a_launches_bench.summary.txt
stamp ncu_list
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tap2_kernel|flash_attn_d64_kernel|gn_fused_kernel|gn_apply_kernel|gn_part_finalize|ln_finalize|temporal_attn" \
  -o $O/a_prof_r02 -f python tools/ncu_target.py all > $O/a_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 $O/a_ncu_full.log
ls -la $O/a_prof_r02.ncu-rep
stamp ncu_full
for wl in ViewCrafter_25_512 ViewCrafter_16; do
  timeout 400 python bench.py --workload $wl --steps 6 --warmup 3 --no-cpu-baseline > $O/a_bench_$wl.json 2> $O/a_bench_$wl.err
  echo "$wl rc=$? $(cut -c1-200 $O/a_bench_$wl.json)"
done
stamp workloads
timeout 500 python bench.py --steps 6 --warmup 3 > $O/a_bench_full.json 2> $O/a_bench_full.err; echo "full bench rc=$? $(cut -c1-160 $O/a_bench_full.json)"
stamp bench_full
timeout 900 python tools/parity_50step.py > $O/a_parity_50step.json 2> $O/a_parity_50step.err; echo "parity50 rc=$?"; head -c 1500 $O/a_parity_50step.json
stamp parity50

#!/bin/bash
# round-2 evidence run (1 GPU): config 2 / config 5 bench lines, 50-step parity, ncu launch list of one bench step, ncu --set full of
# the dominant kernels.  Outputs under gpurun_out/e_*; summaries are copied to profiles/ by hand afterwards.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
# attention A/B: polynomial-exp2 selection patterns (side-by-side builds), default library first
for lib in viewcrafter_b200/libvc_b200.so viewcrafter_b200/libvc_b200_*.so; do
  [ -f "$lib" ] || continue
  VC_B200_LIB=$PWD/$lib timeout 200 python tools/ab_micro.py 2>&1 | grep -E "attn|rror" >> $O/e_ab_attn.txt
done
VC_ATTN_BN64=1 timeout 200 python tools/ab_micro.py 2>&1 | grep -E "attn" | sed "s/^/[force bn64] /" >> $O/e_ab_attn.txt
VC_ATTN_BN64=0 timeout 200 python tools/ab_micro.py 2>&1 | grep -E "attn" | sed "s/^/[force bn128] /" >> $O/e_ab_attn.txt
timeout 200 python tools/ab_micro.py 2>&1 | grep -E "groupnorm|ln_stats" >> $O/e_ab_attn.txt
cat $O/e_ab_attn.txt
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-vae > $O/e_bench_default.json 2> $O/e_bench_default.err; echo "bench default: $(cut -c1-130 $O/e_bench_default.json)"
for wl in ViewCrafter_25_512 ViewCrafter_16; do
  timeout 400 python bench.py --workload $wl --steps 6 --warmup 3 --no-cpu-baseline > $O/e_bench_$wl.json 2> $O/e_bench_$wl.err
  echo "$wl rc=$? $(cut -c1-200 $O/e_bench_$wl.json)"
done
timeout 900 python tools/parity_50step.py > $O/e_parity_50step.json 2> $O/e_parity_50step.err; echo "parity50 rc=$?"; cat $O/e_parity_50step.json | head -40
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_tap|flash_attn|gn_|temporal_attn|ln_|small_linear|im2col|upsample|ddim_|nchw|nhwc|cast_|timestep|peer_" \
  --csv --log-file $O/e_launches_bench.csv python bench.py --steps 1 --warmup 2 --no-graph --no-cpu-baseline --no-gpu-baseline --no-vae > $O/e_launches_bench.log 2>&1
python tools/launch_summary.py $O/e_launches_bench.csv > $O/e_launches_bench.summary.txt 2>&1; head -25 $O/e_launches_bench.summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tap2_kernel|flash_attn_d64_kernel|gn_pipe_kernel|gn_fused_kernel|ln_finalize|temporal_attn" \
  -o $O/e_prof_r02 -f python tools/ncu_target.py all > $O/e_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 $O/e_ncu_full.log
ls -la $O/e_prof_r02.ncu-rep

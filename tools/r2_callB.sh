#!/bin/bash
# round-2 GPU call B (1 GPU): GPU suite, GroupNorm-from-producer micro (v2 normalise pass), GEMM micro with / without the L2 prefetch of the next
# tile's rows, attention micro incl. the ping-pong kernel, bench A/Bs, ncu --set full of the dominant kernels (CSV extracted ON the box: the
# report itself exceeds the 64 MB gpurun_out limit).  Everything printed goes to gpurun_out/b_stdout.txt as well.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
exec > >(tee $O/b_stdout.txt) 2>&1
T0=$(date +%s)
stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi -L
stamp start
# the ping-pong attention kernel first, under a short timeout: if it hangs or fails, every later ping-pong run is skipped
PP_OK=0
VC_ATTN_PP=1 VC_ATTN_BN64=0 timeout 90 python tools/attn_check.py > $O/b_attn_check_pp.txt 2>&1 && PP_OK=1
tail -9 $O/b_attn_check_pp.txt; echo "PP_OK=$PP_OK"
stamp pp_check
timeout 900 python -m pytest tests -m gpu -q -rf --deselect tests/test_multigpu_gpu.py --deselect tests/test_ops_gpu.py::test_attention_kernel_variants > $O/b_pytest.log 2>&1; echo "pytest rc=$?" >> $O/b_pytest.log
tail -25 $O/b_pytest.log
stamp pytest
timeout 200 python tools/gn_parts_micro.py 2>&1 | tee $O/b_gn_parts_micro.txt | tail -30
stamp gn_micro
VC_GEMM_L2PF=1 timeout 200 python tools/bench_gemm.py 2>&1 | sed "s/^/[l2pf=1] /" | tee $O/b_gemm_micro.txt
VC_GEMM_L2PF=0 timeout 200 python tools/bench_gemm.py 2>&1 | sed "s/^/[l2pf=0] /" | tee -a $O/b_gemm_micro.txt
stamp gemm_micro
timeout 200 python tools/ab_micro.py 2>&1 | grep -E "attn|rror" | tee $O/b_ab_attn.txt
VC_ATTN_BN64=1 timeout 90 python tools/attn_check.py 2>&1 | tail -2
VC_ATTN_BN64=0 VC_ATTN_PP=0 timeout 90 python tools/attn_check.py 2>&1 | tail -2
[ $PP_OK = 1 ] && VC_ATTN_PP=1 timeout 120 python tools/ab_micro.py 2>&1 | grep -E "self-attn|rror" | tee -a $O/b_ab_attn.txt
stamp attn_micro
bench() { # name env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-vae > $O/b_bench_$name.json 2> $O/b_bench_$name.err
  echo "bench $name rc=$? $(cut -c1-150 $O/b_bench_$name.json)"; tail -2 $O/b_bench_$name.err
}
bench default VC_NOP=1
bench l2pf0 VC_GEMM_L2PF=0
bench gn0 VC_GN_FROM_PRODUCER=0
bench gnall VC_GN_PARTS_MIN_MB=0
bench gn2 VC_GN_FROM_PRODUCER=2
[ $PP_OK = 1 ] && bench pp VC_ATTN_PP=1
stamp bench_ab
VC_NCU_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tap2_kernel|flash_attn_d64_kernel|gn_fused_kernel|gn_apply_kernel|gn_part_finalize|temporal_attn" \
  -o /tmp/b_prof_r02 -f python tools/ncu_target.py all > $O/b_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 $O/b_ncu_full.log
ls -la /tmp/b_prof_r02.ncu-rep
ncu -i /tmp/b_prof_r02.ncu-rep --page raw --csv > $O/b_ncu_full_raw.csv 2> $O/b_ncu_raw.err; ls -la $O/b_ncu_full_raw.csv
ncu -i /tmp/b_prof_r02.ncu-rep --page details --csv > $O/b_ncu_full_details.csv 2>> $O/b_ncu_raw.err
stamp ncu_full
du -sh $O

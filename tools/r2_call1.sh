#!/bin/bash
# round-2 GPU call 1: full GPU test suite (incl. BASELINE-size parity), bench lines (graph / eager / switches), GN / attention / LN / GEMM
# A/B micro-benchmarks, smoke launch list
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/c1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s -rf > $O/c1_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c1_pytest.log
grep -E "passed|failed|FAILED|ERROR|ViewCrafter_25" $O/c1_pytest.log | tail -40
timeout 900 python bench.py --steps 6 --warmup 3 > $O/c1_bench.json 2> $O/c1_bench.err; echo "bench rc=$?"
tail -c 3500 $O/c1_bench.json; tail -5 $O/c1_bench.err
timeout 300 python bench.py --steps 6 --warmup 3 --no-graph --no-cpu-baseline --no-gpu-baseline --no-vae > $O/c1_bench_nograph.json 2>> $O/c1_bench.err
VC_LN_FROM_PRODUCER=0 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-vae > $O/c1_bench_lnpass.json 2>> $O/c1_bench.err
VC_GN_L2_MB=0 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-vae > $O/c1_bench_gn1launch.json 2>> $O/c1_bench.err
VC_GN_PIPE=0 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-vae > $O/c1_bench_gnchunked.json 2>> $O/c1_bench.err
for f in nograph lnpass gn1launch gnchunked; do echo "$f: $(cut -c1-120 $O/c1_bench_$f.json)"; done
for mb in 0 32 48 64 96; do
  VC_GN_PIPE=0 VC_GN_L2_MB=$mb timeout 200 python tools/ab_micro.py 2>&1 | grep -E "groupnorm" | sed "s/^/[chunked GN_L2_MB=$mb] /" >> $O/c1_ab.txt
done
for mb in 24 48 72; do
  VC_GN_PIPE=1 VC_GN_L2_MB=$mb timeout 200 python tools/ab_micro.py 2>&1 | grep -E "groupnorm" | sed "s/^/[team pipeline GN_L2_MB=$mb] /" >> $O/c1_ab.txt
done
timeout 200 python tools/ab_micro.py 2>&1 | grep -vE "groupnorm" >> $O/c1_ab.txt
for lib in viewcrafter_b200/libvc_b200_*.so; do
  [ -f "$lib" ] || continue
  VC_B200_LIB=$PWD/$lib timeout 200 python tools/ab_micro.py 2>&1 | grep -E "attn|rror" >> $O/c1_ab.txt
done
VC_ATTN_BN64=1 timeout 200 python tools/ab_micro.py 2>&1 | grep -E "attn" >> $O/c1_ab.txt
VC_LN_STATS_UNROLL=1 timeout 200 python tools/ab_micro.py 2>&1 | grep -E "ln_stats" >> $O/c1_ab.txt
timeout 200 python tools/bench_gemm.py > $O/c1_gemm.txt 2>&1
cat $O/c1_ab.txt; cat $O/c1_gemm.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file $O/c1_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > $O/c1_smoke.log 2>&1
python tools/launch_summary.py $O/c1_smoke_launches.csv 2>/dev/null | head -30
tail -3 $O/c1_smoke.log

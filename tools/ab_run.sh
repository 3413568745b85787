# A/B evidence run (gpurun -- 'bash tools/ab_run.sh'): full GPU suite on the default build, then tools/ab_micro.py on the default
# build and on every side-by-side build tools/ab_build.sh left in viewcrafter_b200/, then the step benchmark.
set -x
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/ab_tests.log; cat gpurun_out/ab_tests.log
: > gpurun_out/ab_micro.log
for lib in viewcrafter_b200/libvc_b200.so viewcrafter_b200/libvc_b200_*.so; do
  [ -f "$lib" ] || continue
  VC_B200_LIB=$PWD/$lib timeout 120 python tools/ab_micro.py 2>&1 | grep -E "^\[" >> gpurun_out/ab_micro.log
done
# opt-in experiments selected by environment (same default library)
VC_LN_STATS_UNROLL=1 timeout 120 python tools/ab_micro.py 2>&1 | grep -E "^\[.*ln_stats|Error|error" >> gpurun_out/ab_micro.log
VC_ATTN_BN64=1 timeout 120 python tools/ab_micro.py 2>&1 | grep -E "^\[|Error|error" >> gpurun_out/ab_micro.log
cat gpurun_out/ab_micro.log
timeout 200 python bench.py --steps 4 --warmup 3 2>gpurun_out/ab_bench.err | tail -1 > gpurun_out/ab_bench_default.json; cut -c1-160 gpurun_out/ab_bench_default.json

# A/B evidence run (gpurun -- 'bash tools/ab_run.sh'): full GPU suite on the default build, kernel A/B of the side-by-side builds, step A/B.
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/ab_tests.log; cat gpurun_out/ab_tests.log
for v in "" _base _split _parked _noparked _poly0 _poly2 _poly3; do
  VC_B200_LIB=$PWD/viewcrafter_b200/libvc_b200$v.so timeout 120 python tools/ab_micro.py 2>&1 | grep -E "^\[" >> gpurun_out/ab_micro.log
done
cat gpurun_out/ab_micro.log
timeout 200 python bench.py --steps 4 --warmup 3 2>gpurun_out/ab_bench.err | tail -1 > gpurun_out/ab_bench_default.json; cut -c1-160 gpurun_out/ab_bench_default.json
VC_B200_LIB=$PWD/viewcrafter_b200/libvc_b200_base.so timeout 150 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>>gpurun_out/ab_bench.err | tail -1 > gpurun_out/ab_bench_base.json; cut -c1-160 gpurun_out/ab_bench_base.json

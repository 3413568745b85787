# Round-end evidence run on one B200 (gpurun -- 'bash tools/final_run.sh'): ordered by importance, each leg bounded.
set -x
mkdir -p gpurun_out
K='regex:vc::|gemm_tap|flash_attn|gn_|layernorm|ln_stats|temporal_attn|ddim_|small_linear|im2col|upsample|nchw|ncthw|cast_kernel|timestep|softmax_rows'
timeout 330 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/t_all_r1_final.log; cat gpurun_out/t_all_r1_final.log
timeout 170 python bench.py --steps 4 --warmup 3 2>gpurun_out/bench_final.err | tail -1 > gpurun_out/bench_r1_final.json; cut -c1-200 gpurun_out/bench_r1_final.json
timeout 170 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_final.csv -k "$K" python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches_bench_final.csv
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"gemm_tap|flash_attn|gn_fused|ln_stats|temporal_attn" -c 18 -o gpurun_out/prof_final_r1 -f python tools/ncu_target.py all > gpurun_out/ncu_final.log 2>&1
tail -2 gpurun_out/ncu_final.log
timeout 90 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -3 | tee gpurun_out/smoke_final.log
timeout 120 python tools/bench_vae.py 2>&1 | grep -E "decode|encode" | tee gpurun_out/vae_bench.log

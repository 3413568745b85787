# last-minutes check of the tests that call 2 did not reach (gpurun -- 'bash tools/tail_run.sh')
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
( timeout 70 python -m pytest "tests/test_pipeline_gpu.py::test_ddim_sample_three_steps_vs_oracle" -x -v -p no:cacheprovider 2>&1 | tail -25 ) > gpurun_out/tail_ddim.log; echo "rc=$?" >> gpurun_out/tail_ddim.log; tail -6 gpurun_out/tail_ddim.log
( timeout 60 python -m pytest tests/test_pipeline_gpu.py tests/test_ops_gpu.py -k "resampler or gelu or three_way" -v -p no:cacheprovider 2>&1 | tail -25 ) > gpurun_out/tail_new.log; tail -8 gpurun_out/tail_new.log
( timeout 100 python -m pytest tests/test_unet_gpu.py -x -v -p no:cacheprovider 2>&1 | tail -25 ) > gpurun_out/tail_unet.log; tail -8 gpurun_out/tail_unet.log

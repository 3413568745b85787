#!/bin/bash
# round-2 GPU call 2 (2 GPUs): full GPU suite incl. the 2-GPU frame-sharding checks (peer-memory kernels and NCCL), N=2 bench lines
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/c2_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s -rf > $O/c2_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c2_pytest.log
grep -E "passed|failed|FAILED|ERROR|world|peer exchange|PARALLEL" $O/c2_pytest.log | tail -40
bash tools/r2_multi.sh 2

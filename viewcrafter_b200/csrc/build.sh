#!/bin/bash
# Build libvc_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr ${VC_NVCC_EXTRA:-}"
OUT=../libvc_b200.so
SRCS="host.cu capi.cu gemm_tap.cu gemm_tap2.cu attention.cu temporal_attn.cu norm.cu misc.cu"
mkdir -p build
pids=()
for f in $SRCS; do
  $NVCC $FLAGS -c $f -o build/${f%.cu}.o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o $OUT build/*.o -lcudart
echo "built $(realpath $OUT)"

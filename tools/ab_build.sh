#!/bin/bash
# Build side-by-side variants of libvc_b200.so for A/B runs (here, before gpurun: the .so files travel with the snapshot).
#   bash tools/ab_build.sh "base:-DVC_ATT_SPLIT_KV=0 -DVC_GN_REVERSE=0" "parked:-DVC_ATT_PARKED_WAIT=1" "poly3:-DVC_ATT_POLY_PERIOD=3"
# -> viewcrafter_b200/libvc_b200_<name>.so (git-ignored); load one with VC_B200_LIB=<path>; tools/ab_run.sh runs them all.
set -euo pipefail
cd "$(dirname "$0")/../viewcrafter_b200/csrc"
for v in "$@"; do
  n=${v%%:*}; f=${v#*:}
  ( VC_NVCC_EXTRA="$f" VC_OUT=../libvc_b200_$n.so VC_BUILD_DIR=build_$n bash build.sh 2>&1 | tail -1 ) &
done
wait

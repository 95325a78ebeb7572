#!/bin/bash
# K3 pair kernel at 160 threads: loop form x pairs per thread
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for v in a b c d e; do
  export SSR_B200_LIB=$PWD/build/variants/libssr_b200_$v.so
  echo "== variant '$v'"
  timeout 300 python tools/bench_kernels.py k3 2>&1 | tee gpurun_out/s28_k3_$v.log
done

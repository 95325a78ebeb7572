#!/bin/bash
# ncu of the two K3 kernels after the bank transposition
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_resample_pair -s 1 -c 1 -o gpurun_out/s24_prof_k3_pair python tools/bench_kernels.py k3 > gpurun_out/s24_ncu_k3_pair.log 2>&1
SSR_FORCE_BULK_K3=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_resample_bulk -s 1 -c 1 -o gpurun_out/s24_prof_k3_bulk python tools/bench_kernels.py k3 > gpurun_out/s24_ncu_k3_bulk.log 2>&1
ls -la gpurun_out | grep s24

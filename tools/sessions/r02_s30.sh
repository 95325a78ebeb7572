#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_resample_pair -s 1 -c 1 -o gpurun_out/s30_prof_k3_pair python tools/bench_kernels.py k3 > gpurun_out/s30_ncu_k3_pair.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ssim -s 1 -c 1 -o gpurun_out/s30_prof_k2 python tools/ab_lib.py --flags 15 - > gpurun_out/s30_ncu_k2.log 2>&1
ls -la gpurun_out | grep s30

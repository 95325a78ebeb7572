#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_stft_hard_lowpass_2048 -s 1 -c 1 -o gpurun_out/s40_prof_k4 python tools/bench_kernels.py k4 > gpurun_out/s40_ncu_k4.log 2>&1
ls -la gpurun_out | grep s40

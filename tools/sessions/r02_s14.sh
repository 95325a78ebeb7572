#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stft_splice_istft_2048 -s 1 -c 1 -o gpurun_out/s14_prof_k6 python tools/bench_kernels.py k6 > gpurun_out/s14_ncu_k6.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sosfiltfilt -s 1 -c 1 -o gpurun_out/s14_prof_k7 python tools/bench_kernels.py k7 > gpurun_out/s14_ncu_k7.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stft_hard_lowpass_2048 -s 1 -c 1 -o gpurun_out/s14_prof_k4 python tools/bench_kernels.py k4 > gpurun_out/s14_ncu_k4.log 2>&1
ls -la gpurun_out | grep s14

#!/bin/bash
# Round-end profiling pass (run under gpurun, 1 GPU).  Outputs land in gpurun_out/.
set -x
# 1. launch list of the contract bench command (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu.log 2>&1
# 2. full-set captures, one launch per kernel
ncu --set full --clock-control none --import-source on -k regex:k_stft_metrics_2048 -s 3 -c 1 -o gpurun_out/prof_k1_final \
    python bench.py --pairs 256 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
ncu --set full --clock-control none -k regex:k_ssim -s 1 -c 1 -o gpurun_out/prof_k2 \
    python tools/bench_kernels.py k1 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:k_stft_metrics_pfa -s 1 -c 1 -o gpurun_out/prof_k1_pfa \
    python tools/bench_kernels.py k1blue > /dev/null 2>&1
ncu --set full --clock-control none -k regex:k_resample_tiled -s 1 -c 1 -o gpurun_out/prof_k3 \
    python tools/bench_kernels.py k3 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:k_stft_hard_lowpass_2048 -s 1 -c 1 -o gpurun_out/prof_k4 \
    python tools/bench_kernels.py k4 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep

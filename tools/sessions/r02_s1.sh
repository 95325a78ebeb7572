#!/bin/bash
# round 2, GPU session 1: parity suite, warp-local A/B, bench with the PCM16 e2e path, log_sispec distribution, K3 ncu
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/s1_gpu.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s1_pytest.log 2>&1
tail -5 gpurun_out/s1_pytest.log
timeout 300 python tools/ab_lib.py --k4 - build/variants/libssr_b200_wl.so > gpurun_out/s1_ab_wl.log 2>&1
cat gpurun_out/s1_ab_wl.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err
cat gpurun_out/s1_bench.json; tail -3 gpurun_out/s1_bench.err
timeout 400 python tools/logsispec_distribution.py --pairs 64 --out gpurun_out/s1_logsispec_distribution.json > gpurun_out/s1_logsispec.md 2>&1
cat gpurun_out/s1_logsispec.md
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_resample_tiled -s 1 -c 1 -o gpurun_out/s1_prof_k3 python tools/bench_kernels.py k3 > gpurun_out/s1_ncu_k3.log 2>&1
ls -la gpurun_out/ | tail -15

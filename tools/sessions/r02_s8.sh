#!/bin/bash
# round 2, GPU session 8: parity suite, fresh source-level ncu captures of K1 (2048), K2 (SSIM), PFA
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/s8_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/s8_pytest.log; grep -E "^E  |^FAILED" gpurun_out/s8_pytest.log | cut -c1-300 | head -20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stft_metrics_2048 -s 3 -c 1 -o gpurun_out/s8_prof_k1 \
    python bench.py --pairs 256 --steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-oracle-check > gpurun_out/s8_ncu_k1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ssim -s 1 -c 1 -o gpurun_out/s8_prof_k2 \
    python tools/bench_kernels.py k1 > gpurun_out/s8_ncu_k2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stft_metrics_pfa -s 1 -c 1 -o gpurun_out/s8_prof_pfa \
    python tools/bench_kernels.py k1blue > gpurun_out/s8_ncu_pfa.log 2>&1
ls -la gpurun_out | grep s8

#!/bin/bash
# round 2, GPU session 7: full parity suite, host CPU description, ncu of the TMA-staged K3 and of K8
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
lscpu | grep -E "Model name|^CPU\(s\)|Flags|NUMA|Thread|Socket" | cut -c1-400 > gpurun_out/s7_lscpu.txt
( time timeout 1500 python -m pytest tests -m gpu -q -s ) > gpurun_out/s7_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/s7_pytest.log; grep -E "^E  |^FAILED|bit-identical|LSD reference|^dense|^fft|resample " gpurun_out/s7_pytest.log | cut -c1-400 | head -40
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_resample_bulk -s 1 -c 1 -o gpurun_out/s7_prof_k3 python tools/bench_kernels.py k3 > gpurun_out/s7_ncu_k3.log 2>&1
ls -la gpurun_out | grep s7

#!/bin/bash
# K3 pair kernel: compact tap table, 16 vs 8 pairs per thread
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q -k "resampl or k3 or polyphase or lowpass or helper" ) > gpurun_out/s26_pytest.log 2>&1
tail -3 gpurun_out/s26_pytest.log; grep -E "^E  |^FAILED" gpurun_out/s26_pytest.log | cut -c1-300 | head -20
timeout 300 python tools/bench_kernels.py k3 > gpurun_out/s26_k3_rp16.log 2>&1; cat gpurun_out/s26_k3_rp16.log
SSR_B200_LIB=$PWD/build/variants/libssr_b200_k3rp8.so timeout 300 python tools/bench_kernels.py k3 > gpurun_out/s26_k3_rp8.log 2>&1; cat gpurun_out/s26_k3_rp8.log

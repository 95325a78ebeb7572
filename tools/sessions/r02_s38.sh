#!/bin/bash
# after the K3 table refactor: resampler parity, misaligned-workspace test, K3 timing
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q -k "resampl or misaligned or helper or lowpass or load" ) > gpurun_out/s38_pytest.log 2>&1
tail -3 gpurun_out/s38_pytest.log; grep -E "^E  |^FAILED" gpurun_out/s38_pytest.log | cut -c1-300 | head
timeout 300 python tools/bench_kernels.py k3 2>&1 | tee gpurun_out/s38_k3.log

#!/bin/bash
# K3 pair kernel: parity of the resampler tests, timing against the bulk kernel
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q -k "resampl or k3 or polyphase or lowpass or helper" ) > gpurun_out/s23_pytest.log 2>&1
tail -5 gpurun_out/s23_pytest.log; grep -E "^E  |^FAILED" gpurun_out/s23_pytest.log | cut -c1-300 | head -20
timeout 300 python tools/bench_kernels.py k3 > gpurun_out/s23_k3_pair.log 2>&1; cat gpurun_out/s23_k3_pair.log
SSR_FORCE_BULK_K3=1 timeout 300 python tools/bench_kernels.py k3 > gpurun_out/s23_k3_bulk.log 2>&1; cat gpurun_out/s23_k3_bulk.log

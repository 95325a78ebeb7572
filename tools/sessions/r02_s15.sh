#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python tools/bench_kernels.py k6 k7 > gpurun_out/s15_k67.log 2>&1; cat gpurun_out/s15_k67.log
timeout 900 python -m pytest tests -m gpu -q -k "sosfiltfilt or postprocessing or helper or goldens or float64" > gpurun_out/s15_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/s15_pytest.log; grep -E "^E  |^FAILED" gpurun_out/s15_pytest.log | cut -c1-300 | head

#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python tools/bench_kernels.py k1_f64est > gpurun_out/s17_f64est.log 2>&1; cat gpurun_out/s17_f64est.log
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/s17_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/s17_pytest.log; grep -E "^E  |^FAILED" gpurun_out/s17_pytest.log | cut -c1-300 | head -20

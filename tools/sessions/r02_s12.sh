#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python tools/bench_kernels.py k4 > gpurun_out/s12_k4.log 2>&1; cat gpurun_out/s12_k4.log
timeout 900 python -m pytest tests -m gpu -q -k "lowpass or helper or many_small or k4 or smoke" > gpurun_out/s12_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/s12_pytest.log; grep -E "^E  |^FAILED" gpurun_out/s12_pytest.log | cut -c1-300 | head
timeout 600 python tools/bench_configs.py cfg3 cfg4 > gpurun_out/s12_configs.log 2>&1; cat gpurun_out/s12_configs.log

#!/bin/bash
# round 2, GPU session 9: parity suite after the test fixes, cfg5 (strong scaling) at N = 1, kernel timings
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/s9_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/s9_pytest.log; grep -E "^E  |^FAILED" gpurun_out/s9_pytest.log | cut -c1-300 | head -20
timeout 600 python bench.py --config cfg5 --steps 5 --warmup 3 > gpurun_out/s9_cfg5_n1.json 2> gpurun_out/s9_cfg5_n1.err
cat gpurun_out/s9_cfg5_n1.json; tail -3 gpurun_out/s9_cfg5_n1.err
timeout 900 python tools/bench_kernels.py > gpurun_out/s9_kernel_timings.log 2>&1
cat gpurun_out/s9_kernel_timings.log
timeout 900 python tools/bench_configs.py cfg3 cfg4 > gpurun_out/s9_configs.log 2>&1
cat gpurun_out/s9_configs.log

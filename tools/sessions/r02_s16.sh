#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python tools/bench_kernels.py k1_f64est > gpurun_out/s16_f64est.log 2>&1; cat gpurun_out/s16_f64est.log

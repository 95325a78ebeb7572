#!/bin/bash
# sanitizers on the final kernels (k_resample_pair, interleaved K1 -> K2 image, ...)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/s31_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -4 gpurun_out/s31_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_small.py quick > gpurun_out/s31_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -4 gpurun_out/s31_racecheck.log

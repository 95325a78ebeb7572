#!/bin/bash
# round 2, GPU session 10: timings of the new kernels (K0, K4d, K8, K1 at hop 441), bench, sanitizers on the final build
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python tools/bench_kernels.py k0 k1_441 k4d k8 > gpurun_out/s10_kernel_timings_new.log 2>&1
cat gpurun_out/s10_kernel_timings_new.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s10_bench.json 2> gpurun_out/s10_bench.err
cat gpurun_out/s10_bench.json; tail -3 gpurun_out/s10_bench.err
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py quick > gpurun_out/s10_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/s10_memcheck.log
tail -6 gpurun_out/s10_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_small.py quick > gpurun_out/s10_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/s10_racecheck.log
tail -6 gpurun_out/s10_racecheck.log

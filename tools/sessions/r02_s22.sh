#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s22_bench.json 2> gpurun_out/s22_bench.err
cat gpurun_out/s22_bench.json; tail -3 gpurun_out/s22_bench.err

#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/ab_lib.py --flags 15 - build/variants/libssr_b200_copyloop.so build/variants/libssr_b200_ringmov.so - build/variants/libssr_b200_copyloop.so build/variants/libssr_b200_ringmov.so > gpurun_out/s34_ab.log 2>&1; cat gpurun_out/s34_ab.log

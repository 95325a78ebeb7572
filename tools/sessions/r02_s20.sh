#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python tools/ab_lib.py --hop 441 --flags 1,7,15 - build/variants/libssr_b200_tw2t.so > gpurun_out/s20_ab441.log 2>&1; cat gpurun_out/s20_ab441.log
timeout 600 python tools/ab_lib.py --flags 1,7,15 - build/variants/libssr_b200_tw2tnr.so > gpurun_out/s20_ab512.log 2>&1; cat gpurun_out/s20_ab512.log

#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/ab_lib.py --nfft 2229 --hop 480 --flags 1,15 --pairs 256 - build/variants/libssr_b200_cwinlate.so - build/variants/libssr_b200_cwinlate.so > gpurun_out/s36_ab_pfa.log 2>&1; cat gpurun_out/s36_ab_pfa.log

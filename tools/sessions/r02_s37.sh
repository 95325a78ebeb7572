#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
V=build/variants
timeout 400 python tools/ab_lib.py --nfft 2229 --hop 480 --flags 1,15 --pairs 256 - $V/libssr_b200_cwinlate.so $V/libssr_b200_alllate.so $V/libssr_b200_alllate_post.so $V/libssr_b200_cwinlate_post.so > gpurun_out/s37_ab_pfa.log 2>&1; cat gpurun_out/s37_ab_pfa.log

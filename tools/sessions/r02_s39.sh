#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 300 python -m pytest tests -m gpu -q -k "misaligned" ) 2>&1 | tail -2
timeout 300 python tools/ab_lib.py --nfft 2229 --hop 480 --flags 1,15 --pairs 256 - build/variants/libssr_b200_pfa2.so - build/variants/libssr_b200_pfa2.so > gpurun_out/s39_ab_pfa.log 2>&1; cat gpurun_out/s39_ab_pfa.log

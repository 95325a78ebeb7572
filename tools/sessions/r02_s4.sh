#!/bin/bash
# round 2, GPU session 4: A/B of K1 (item pipeline, sample loads at frame start), K4 / PFA (twiddle prefetch, warp-local);
# dense stft_hard mode tests
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python tools/ab_lib.py --k4 - build/variants/libssr_b200_prereg.so build/variants/libssr_b200_wlbase.so > gpurun_out/s4_ab.log 2>&1
cat gpurun_out/s4_ab.log
timeout 600 python tools/ab_lib.py --nfft 2229 --hop 480 --pairs 256 --flags 1,15 - build/variants/libssr_b200_wlbase.so > gpurun_out/s4_ab_pfa.log 2>&1
cat gpurun_out/s4_ab_pfa.log
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/s4_pytest.log 2>&1
tail -30 gpurun_out/s4_pytest.log

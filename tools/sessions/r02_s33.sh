#!/bin/bash
# PFA: prefetch moved behind the pass-1 stores (A/B), K1<15> copy-out, K2 in-place ring: parity + timing
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x -k "not dense and not resampl" ) > gpurun_out/s33_pytest.log 2>&1
tail -3 gpurun_out/s33_pytest.log; grep -E "^E  |^FAILED" gpurun_out/s33_pytest.log | cut -c1-300 | head
timeout 300 python tools/ab_lib.py --nfft 2229 --hop 480 --flags 1,15 --pairs 256 - build/variants/libssr_b200_pfa_early.so - build/variants/libssr_b200_pfa_early.so > gpurun_out/s33_ab_pfa.log 2>&1; cat gpurun_out/s33_ab_pfa.log
timeout 300 python tools/ab_lib.py --flags 1,7,15 - > gpurun_out/s33_ab.log 2>&1; cat gpurun_out/s33_ab.log

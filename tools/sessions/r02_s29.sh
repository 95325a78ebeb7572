#!/bin/bash
# interleaved K1 -> K2 spectrogram image, 16-byte row copies in K2: full parity suite + all-four timing
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/s29_pytest.log 2>&1
tail -4 gpurun_out/s29_pytest.log; grep -E "^E  |^FAILED" gpurun_out/s29_pytest.log | cut -c1-300 | head -20
timeout 300 python tools/ab_lib.py --flags 1,7,15 - > gpurun_out/s29_ab.log 2>&1; cat gpurun_out/s29_ab.log
timeout 300 python tools/ab_lib.py --nfft 2229 --hop 480 --flags 1,15 --pairs 256 - > gpurun_out/s29_ab_pfa.log 2>&1; cat gpurun_out/s29_ab_pfa.log

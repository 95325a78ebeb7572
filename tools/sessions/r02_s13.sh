#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python tools/ab_lib.py --flags 15,8 - build/variants/libssr_b200_k2old.so > gpurun_out/s13_ab.log 2>&1; cat gpurun_out/s13_ab.log
timeout 900 python -m pytest tests -m gpu -q -k "ragged_batch or goldens_through or flag_subsets or minimal or poisoning or unusual or fuzz" > gpurun_out/s13_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/s13_pytest.log; grep -E "^E  |^FAILED" gpurun_out/s13_pytest.log | cut -c1-300 | head

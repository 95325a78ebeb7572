#!/bin/bash
# round 2, GPU session 6: full parity suite (dense mode alignment fix, K3 bulk, K8 alignment), bench, memcheck of the new kernels
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/s6_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/s6_pytest.log; grep -E "^E  |^FAILED" gpurun_out/s6_pytest.log | head -30
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s6_bench.json 2> gpurun_out/s6_bench.err
cat gpurun_out/s6_bench.json; tail -3 gpurun_out/s6_bench.err
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "xcorr or bulk_staged or dense_stft or pcm16" > gpurun_out/s6_memcheck.log 2>&1
tail -15 gpurun_out/s6_memcheck.log

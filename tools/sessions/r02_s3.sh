#!/bin/bash
# round 2, GPU session 3: K1 A/B (pass-2 twiddle prefetch, window loads ahead), parity suite, bench
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python tools/ab_lib.py - build/variants/libssr_b200_wlbase.so build/variants/libssr_b200_w0.so build/variants/libssr_b200_w16.so build/variants/libssr_b200_serial.so > gpurun_out/s3_ab.log 2>&1
cat gpurun_out/s3_ab.log
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/s3_pytest.log 2>&1
tail -8 gpurun_out/s3_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err
cat gpurun_out/s3_bench.json; tail -3 gpurun_out/s3_bench.err

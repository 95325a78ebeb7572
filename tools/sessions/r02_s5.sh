#!/bin/bash
# round 2, GPU session 5: K1 item-pipeline ablation, PFA prefetch/warp-local ablation, K3 bulk (TMA-staged) kernel, tests
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python tools/ab_lib.py --k4 - build/variants/libssr_b200_noahead.so build/variants/libssr_b200_nonextpf.so build/variants/libssr_b200_nowl.so > gpurun_out/s5_ab.log 2>&1
cat gpurun_out/s5_ab.log
timeout 600 python tools/ab_lib.py --nfft 2229 --hop 480 --pairs 256 --flags 1,15 - build/variants/libssr_b200_noahead.so build/variants/libssr_b200_nowl.so > gpurun_out/s5_ab_pfa.log 2>&1
cat gpurun_out/s5_ab_pfa.log
SSR_FORCE_OLD_K3=1 timeout 300 python tools/bench_kernels.py k3 > gpurun_out/s5_k3_old.log 2>&1
timeout 300 python tools/bench_kernels.py k3 > gpurun_out/s5_k3_new.log 2>&1
cat gpurun_out/s5_k3_old.log gpurun_out/s5_k3_new.log
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/s5_pytest.log 2>&1
tail -30 gpurun_out/s5_pytest.log

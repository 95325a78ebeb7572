#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_stft_metrics_2048 -s 2 -c 1 -o gpurun_out/s32_prof_k1_15 python tools/ab_lib.py --flags 15 --pairs 256 - > gpurun_out/s32_ncu_k1_15.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_stft_metrics_pfa -s 2 -c 1 -o gpurun_out/s32_prof_pfa python tools/ab_lib.py --nfft 2229 --hop 480 --flags 1 --pairs 256 - > gpurun_out/s32_ncu_pfa.log 2>&1
ls -la gpurun_out | grep s32

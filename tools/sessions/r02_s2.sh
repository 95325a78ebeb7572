#!/bin/bash
# round 2, GPU session 2: rest of the parity suite, bench, ncu source-level capture of K1 (warp-local build)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/s2_pytest.log 2>&1
tail -8 gpurun_out/s2_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err
cat gpurun_out/s2_bench.json; tail -3 gpurun_out/s2_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stft_metrics_2048 -s 3 -c 1 -o gpurun_out/s2_prof_k1 \
    python bench.py --pairs 256 --steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-oracle-check > gpurun_out/s2_ncu_k1.log 2>&1
ls -la gpurun_out/ | tail -8

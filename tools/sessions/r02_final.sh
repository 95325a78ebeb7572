#!/bin/bash
# round 2, final evidence pass: parity suite, contract bench, kernel / config timings, launch list under ncu
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/final_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/final_pytest.log; grep -E "^E  |^FAILED" gpurun_out/final_pytest.log | cut -c1-300 | head -20
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
cat gpurun_out/final_bench.json; tail -3 gpurun_out/final_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
cat gpurun_out/final_bench_reference.json
timeout 900 python tools/bench_kernels.py > gpurun_out/final_kernel_timings.log 2>&1
cat gpurun_out/final_kernel_timings.log
timeout 900 python tools/bench_configs.py cfg3 cfg4 > gpurun_out/final_configs.log 2>&1; cat gpurun_out/final_configs.log
timeout 600 python bench.py --config cfg5 --steps 5 --warmup 3 > gpurun_out/final_cfg5_n1.json 2>/dev/null; cat gpurun_out/final_cfg5_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/final_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-oracle-check > gpurun_out/final_bench_under_ncu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log

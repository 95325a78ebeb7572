#!/bin/bash
# round 2, multi-GPU session: contract bench (weak scaling, device-resident + e2e) and cfg5 (strong scaling) at N ranks
# usage: bash tools/sessions/r02_mgpu.sh N
N=${1:-2}
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/mgpu_topo_n$N.txt 2>&1
P=29517
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/mgpu_bench_n$N.json 2> gpurun_out/mgpu_bench_n$N.err
cat gpurun_out/mgpu_bench_n$N.json; tail -3 gpurun_out/mgpu_bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+1)) \
    bench.py --config cfg5 --gpus $N --steps 5 --warmup 3 > gpurun_out/mgpu_cfg5_n$N.json 2> gpurun_out/mgpu_cfg5_n$N.err
cat gpurun_out/mgpu_cfg5_n$N.json; tail -3 gpurun_out/mgpu_cfg5_n$N.err

#!/bin/bash
N=${1:-4}
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/h2d_probe.py > gpurun_out/h2d_probe_n$N.json 2> gpurun_out/h2d_probe_n$N.err
cat gpurun_out/h2d_probe_n$N.json; tail -2 gpurun_out/h2d_probe_n$N.err

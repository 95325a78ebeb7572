#!/bin/bash
# K3 pair kernel: batched window loads; CTA size / register cap variants
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q -k "resampl or k3 or polyphase or lowpass or helper" ) > gpurun_out/s27_pytest.log 2>&1
tail -3 gpurun_out/s27_pytest.log; grep -E "^E  |^FAILED" gpurun_out/s27_pytest.log | cut -c1-300 | head -20
for v in "" k3tp160 k3minb2; do
  if [ -z "$v" ]; then unset SSR_B200_LIB; else export SSR_B200_LIB=$PWD/build/variants/libssr_b200_$v.so; fi
  echo "== variant '$v'"
  timeout 300 python tools/bench_kernels.py k3 2>&1 | tee gpurun_out/s27_k3_$v.log
done

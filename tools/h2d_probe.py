#!/usr/bin/env python
"""Host-to-device copy bandwidth of every rank at once (pinned memory, cudaMemcpyAsync on one stream per rank):
names the limiter of the end-to-end numbers of bench.py at N > 1.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_probe.py

Prints one JSON line: per-rank GB/s when all ranks copy concurrently, their sum, and rank 0 copying alone."""
import json
import os
import time

import torch
import torch.distributed as td

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    td.init_process_group("nccl", device_id=dev)
n = 512 << 20  # 512 MiB per copy
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h.fill_(1)
d = torch.empty(n, dtype=torch.uint8, device=dev)


def run(iters):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    return iters * n / (time.perf_counter() - t0) / 1e9


run(2)
if world > 1:
    td.barrier()
together = run(8)
t = torch.tensor([together], dtype=torch.float64, device=dev)
allv = [torch.zeros_like(t) for _ in range(world)]
if world > 1:
    td.all_gather(allv, t)
    td.barrier()
else:
    allv = [t]
alone = run(8) if rank == 0 else 0.0
if world > 1:
    td.barrier()
if rank == 0:
    per = [round(float(v.item()), 1) for v in allv]
    print(json.dumps({"n_gpus": world, "copy_bytes": n, "h2d_GBps_per_rank_concurrent": per,
                      "h2d_GBps_sum_concurrent": round(sum(per), 1), "h2d_GBps_rank0_alone": round(alone, 1),
                      "affinity_cpus": len(os.sched_getaffinity(0))}), flush=True)
if world > 1:
    td.destroy_process_group()

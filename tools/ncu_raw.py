#!/usr/bin/env python
"""Print selected raw metrics of the first kernel in an .ncu-rep (reads `ncu -i REP --page raw --csv`)."""
import csv, subprocess, sys
rep = sys.argv[1]
pats = sys.argv[2:] or ["gpu__time_duration.sum", "smsp__issue_active.avg.pct", "lsu_wavefronts", "pipe_fma_cycles_active.avg.pct",
                         "smsp__inst_executed.sum", "registers_per_thread", "warps_active.avg.pct", "bank_conflicts_pipe_lsu_mem_shared.sum",
                         "dram__bytes_read.sum", "dram__bytes_write.sum", "occupancy_limit", "issue_stalled.*per_issue_active", "lts__t_bytes.sum",
                         "l1tex__t_sector_hit_rate", "lts__t_sector_hit_rate", "dram__throughput.avg.pct", "l1tex__m_xbar2l1tex_read_bytes.sum"]
import re
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
h, u, v = r[0], r[1], r[2]
for i, name in enumerate(h):
    if any(re.search(p, name) for p in pats):
        print(f"{name} [{u[i]}] = {v[i]}")

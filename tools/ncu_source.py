#!/usr/bin/env python
"""Summarise the SASS source page of an .ncu-rep (dev tool): opcode histogram by executed warp-instructions, the lines
with the most stall samples (with their neighbours) and, with --regions, samples between user-given SASS line numbers.

    python tools/ncu_source.py REP [--top 25] [--ctx 2]
"""
import argparse, csv, subprocess, collections
ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("--top", type=int, default=25)
ap.add_argument("--ctx", type=int, default=1)
ap.add_argument("--dump", action="store_true", help="print every line: index, samples, executed, SASS")
a = ap.parse_args()
out = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr_i]
iS, iSamp, iExec = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
L = [(r[iS].strip(), int(r[iSamp] or 0), int(r[iExec] or 0)) for r in rows[hdr_i + 1:] if len(r) > iExec]
tot_s, tot_e = sum(x[1] for x in L), sum(x[2] for x in L)
print(f"lines {len(L)}  stall samples {tot_s}  warp-instructions {tot_e}")
hist = collections.Counter()
for src, s, e in L:
    op = src.split()[1] if src.startswith("@") else src.split()[0]
    hist[op.split(".")[0]] += e
print("opcode histogram (share of executed warp-instructions):")
print("  " + ", ".join(f"{k} {v / tot_e * 100:.1f}%" for k, v in hist.most_common(24)))
if a.dump:
    for i, (src, s, e) in enumerate(L):
        print(f"{i:5d} {s:7d} {e:10d}  {src}")
else:
    order = sorted(range(len(L)), key=lambda i: -L[i][1])[:a.top]
    print(f"top {a.top} lines by stall samples:")
    for i in order:
        lo, hi = max(0, i - a.ctx), min(len(L), i + a.ctx + 1)
        for j in range(lo, hi):
            mark = ">>" if j == i else "  "
            print(f"{mark}{j:5d} {L[j][1] / tot_s * 100:6.2f}% {L[j][2]:10d}  {L[j][0][:110]}")
        print()

#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel from an ncu report (needs -lineinfo and
--import-source on): scripts/ncu_hot_lines.py REPORT.ncu-rep [TOP]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = next(r for r in rows if r and r[0] == "Line No")
ie, isamp, ith = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
L, tot, tots, fn = [], 0, 0, ""
for r in rows:
    if r and r[0] == "File Path":
        fn = r[1].split("/")[-1]
    if len(r) > ith and r[0].isdigit() and r[2] == "-":
        n, s, t = int(r[ie]), int(r[isamp]), int(r[ith])
        tot += n; tots += s
        L.append((n, s, t, fn, r[0], r[1].strip()[:100]))
print(f"total warp instructions {tot}, samples {tots}")
for n, s, t, f, l, src in sorted(L, reverse=True)[:top]:
    print(f"{n:>11} {100*n/tot:5.1f}%  smp {100*s/max(tots,1):5.1f}%  thr/inst {t/max(n,1):5.1f}  {f}:{l}: {src}")

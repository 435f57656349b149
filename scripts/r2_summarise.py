#!/usr/bin/env python
"""Runs on the GPU box after scripts/r2_profile.sh captured the reports: turns every
gpurun_out/*.ncu-rep into a small JSON (scripts/ncu_summary.py) + a hot-line listing, so only
kilobytes travel back (gpurun returns at most 64 MiB).  Move counts come from the bench line the
profiled command itself printed."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")


def bench_line(log):
    if not os.path.exists(os.path.join(OUT, log)):
        return {}
    for ln in reversed(open(os.path.join(OUT, log)).read().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)
    return {}


def summarise(rep, out, kernel, derived):
    rep = os.path.join(OUT, rep)
    if not os.path.exists(rep):
        print("missing", rep)
        return
    args = [sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep, os.path.join(OUT, out),
            "--kernel", kernel, "--derived"] + [f"{k}={v}" for k, v in derived.items()]
    subprocess.run(args, check=False, stdout=subprocess.DEVNULL)
    hot = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_hot_lines.py"), rep, "400"],
                         capture_output=True, text=True).stdout
    open(os.path.join(OUT, out.replace(".json", "_hot_lines.txt")), "w").write(hot)


summarise("r2_nq_step_v2.ncu-rep", "r2_ncu_full_nq_step_kernel_v2.json", "nq_step_kernel_v2",
          {"moves": 296 * 49995000, "chains_in_capture": 296, "workload": "n=10000,296_chains,1_step"})
for wl in ("es50", "es2000", "es50x3", "es2000x3"):
    d = bench_line(f"r2_prof_{wl}.log")
    moves = int(d.get("moves_scored_timed", 0))
    summarise(f"r2_es_step_{wl}.ncu-rep", f"r2_ncu_full_es_step_kernel_{wl}.json", "es_step_kernel",
              {"moves": moves, "workload": wl + ",one_launch_of_64_chain_steps"})
summarise("r2_nqb_scan_n1m.ncu-rep", "r2_ncu_full_nqb_scan_packed_kernel_n1m.json", "nqb_scan_packed_kernel",
          {"moves": 499999500000, "n": 1000000, "workload": "n=1000000,one_full_neighbourhood_step"})
for f in os.listdir(OUT):  # the reports themselves stay on the box
    if f.endswith(".ncu-rep") and f != "r2_nq_step_v2.ncu-rep":
        os.remove(os.path.join(OUT, f))

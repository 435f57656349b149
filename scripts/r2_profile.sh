#!/bin/bash
# Round-2 evidence run (one GPU box): launch list of the default bench + one `ncu --set full`
# capture per dominant kernel.  Every profiled command has already exited 0 without ncu.
# Reports land in gpurun_out/; scripts/ncu_summary.py turns them into profiles/r2_*.json here.
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
B="python bench.py --no-cpu-baseline"
# 1. launch list of the default command (kernel SHARE of the step)
$B --steps 2 --warmup 1 > gpurun_out/r2_prof_plain.log 2>&1 || exit 1
$NCU --metrics gpu__time_duration.sum -c 4000 --csv --log-file gpurun_out/r2_launches_bench_default.csv \
    $B --steps 2 --warmup 1 > gpurun_out/r2_prof_launches.log 2>&1
# 2. dominant kernel, 296 chains = two waves of 148 CTAs (the full 4096-chain launch x ~40 replays is minutes)
$NCU --set full --import-source on -k regex:nq_step_kernel_v2 -s 1 -c 1 -f -o gpurun_out/r2_nq_step_v2 \
    $B --chains 296 --steps 1 --warmup 1 --no-secondary --no-e2e > gpurun_out/r2_prof_nq.log 2>&1
# 3. scheduling kernels: the timed 64-step launch of each workload
for wl in es50 es2000 es50x3 es2000x3; do
  $NCU --set full --import-source on -k regex:es_step_kernel -s 1 -c 1 -f -o gpurun_out/r2_es_step_$wl \
      $B --workload $wl --steps 1 --no-e2e > gpurun_out/r2_prof_$wl.log 2>&1
done
# 4. one n = 10^6 board: the packed global scan of one full-neighbourhood step
$NCU --set full --import-source on -k regex:nqb_scan_packed_kernel -s 1 -c 1 -f -o gpurun_out/r2_nqb_scan_n1m \
    $B --workload nq1m --steps 1 --warmup 1 --no-e2e > gpurun_out/r2_prof_nq1m.log 2>&1
python scripts/r2_summarise.py
ls -la gpurun_out/

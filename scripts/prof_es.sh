#!/bin/bash
# full captures of the scheduling step kernel (es50, es2000) with per-line hot spots
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e"
for wl in "$@"; do
  ncu --clock-control none --set full --import-source on -k regex:es_step_kernel -s 1 -c 1 -f -o gpurun_out/p_$wl \
      $B --workload $wl --steps 1 > gpurun_out/p_$wl.log 2>&1
  python scripts/ncu_hot_lines.py gpurun_out/p_$wl.ncu-rep 400 > gpurun_out/p_${wl}_hot.txt
  python scripts/ncu_summary.py gpurun_out/p_$wl.ncu-rep gpurun_out/p_$wl.json --kernel es_step_kernel > /dev/null
  ncu -i gpurun_out/p_$wl.ncu-rep --page details --csv 2>/dev/null | grep -i -E "stall|Warp Cycles Per|Issue Slot|Eligible|Active Warps" | head -40 > gpurun_out/p_${wl}_stalls.txt
  rm -f gpurun_out/p_$wl.ncu-rep
done

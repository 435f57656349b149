#!/bin/bash
# n = 10^6 board: persisting-L2 window on / off (and optional alternative builds via CS_B200_LIB), interleaved
for rep in 1 2; do
  for p in 1 0; do
    CS_NQB_L2_PERSIST=$p python bench.py --workload nq1m --steps 3 --warmup 1 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nq1m persist=$p', '%.4g moves/s'%d['value'], '%.2f ms/step'%d['ms_per_step'])"
  done
done

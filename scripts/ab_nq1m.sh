#!/bin/bash
# n = 10^6 board: current build against alternative builds of the same library (CS_B200_LIB), interleaved
for lib in "" "$@" "" "$@"; do
  CS_B200_LIB=$lib python bench.py --workload nq1m --steps 3 --warmup 1 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nq1m lib=${lib##*/}', '%.4g moves/s'%d['value'], '%.2f ms/step'%d['ms_per_step'])"
done

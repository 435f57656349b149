#!/bin/bash
# packed day sets (rotas of <= 38 days): device-timed rates with the knob off / on (parity: the es suites)
run() { python bench.py --workload $1 --steps 8 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'packed=$2', '%.4g moves/s'%d['value'], '%.4f ms/step'%d['ms_per_step'])"; }
for rep in 1 2; do for wl in es50 es50x3; do
  CS_ES_NO_PACKED_DAY_SETS=1 run $wl off
  env -u CS_ES_NO_PACKED_DAY_SETS bash -c "$(declare -f run); run $wl on"
done; done

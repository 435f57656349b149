#!/bin/bash
# CTA-size sweep of the scheduling kernels through the CS_ES_THREADS knob (device-timed rates)
for spec in "es50x3 64" "es50x3 96" "es50x3 128" "es50x3 192" "es50x3 256" "es2000x3 256" "es2000x3 384" "es2000x3 448" "es2000x3 512" "es50 32" "es50 64" "es2000 96" "es2000 128" "es2000 192"; do
  set -- $spec
  CS_ES_THREADS=$2 python bench.py --workload $1 --steps 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 threads=$2', '%.4g moves/s'%d['value'], '%.4f ms/step'%d['ms_per_step'])"
done

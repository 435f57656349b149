#!/bin/bash
# A/B of scheduling-kernel build variants (alternative builds of the same library via CS_B200_LIB)
run() {  # workload, lib, launches
  for rep in 1 2; do
    CS_B200_LIB=$2 python bench.py --workload $1 --steps ${3:-8} --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'lib=${2##*/}', '%.4g moves/s'%d['value'], '%.4f ms/step'%d['ms_per_step'])"
  done
}
for lib in "" "$@"; do run es2000 "$lib"; run es50 "$lib"; done

#!/bin/bash
# A/B of headline-kernel build variants (alternative builds of the same library via CS_B200_LIB)
python -m pytest tests/test_nq_packed_gpu.py -q -x -k trajectory 2>&1 | tail -1
for lib in "" "$@"; do
  for rep in 1 2; do
    CS_B200_LIB=$lib python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lib=${lib##*/}', '%.5g moves/s'%d['value'], '%.3f ms/step'%d['ms_per_step'])"
  done
done

#!/bin/bash
# scheduling kernels: parity first, then device-timed rates (optionally against alternative builds via CS_B200_LIB)
python -m pytest tests/test_es_gpu.py tests/test_fuzz_gpu.py tests/test_es_reference_mode_gpu.py tests/test_es_slots_gpu.py tests/test_ils_gpu.py -q -x -k "not nqueens and not nq_" 2>&1 | tail -1
run() {  # workload, lib, launches
  for rep in 1 2; do
    CS_B200_LIB=$2 python bench.py --workload $1 --steps ${3:-8} --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'lib=${2##*/}', '%.4g moves/s'%d['value'], '%.4f ms/step'%d['ms_per_step'])"
  done
}
WLS=${WLS:-es2000 es50 es50x3 es2000x3}
for lib in "" "$@"; do for wl in $WLS; do run $wl "$lib" $([ $wl = es2000x3 ] && echo 2 || ([ $wl = es50x3 ] && echo 4 || echo 8)); done; done

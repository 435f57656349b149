#!/bin/bash
# A/B of scheduling-kernel variants on one GPU box (parity first, then device-timed rates)
mkdir -p gpurun_out
python -m pytest tests/test_es_gpu.py tests/test_fuzz_gpu.py tests/test_es_reference_mode_gpu.py tests/test_es_slots_gpu.py tests/test_ils_gpu.py -q -x -k "not nqueens and not nq_" > gpurun_out/ab_es_tests.log 2>&1
tail -2 gpurun_out/ab_es_tests.log
run() {  # workload, env assignment, launches
  for rep in 1 2; do
    env $2 python bench.py --workload $1 --steps ${3:-8} --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', '$2', '%.4g moves/s'%d['value'], '%.4f ms/step'%d['ms_per_step'])"
  done
}
run es2000 X=1
run es50 X=1
run es50x3 X=1 4
run es2000x3 X=1 2

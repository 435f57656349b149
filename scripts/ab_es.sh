#!/bin/bash
# A/B of scheduling-kernel variants on one GPU box (parity first, then device-timed rates)
mkdir -p gpurun_out
python -m pytest tests/test_es_gpu.py tests/test_fuzz_gpu.py tests/test_es_reference_mode_gpu.py -q -x -k "not nqueens" > gpurun_out/ab_es_tests.log 2>&1
tail -2 gpurun_out/ab_es_tests.log
for lib in "" "$PWD/gpurun_out_libs_bq4.so"; do
  for wl in es2000 es50; do
    for rep in 1 2; do
      CS_B200_LIB=$lib python bench.py --workload $wl --steps 8 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl', 'lib=${lib##*/}', '%.4g moves/s'%d['value'], '%.4f ms/step'%d['ms_per_step'], 'kernel %.4f'%d['kernel_ms_per_step'])"
    done
  done
done

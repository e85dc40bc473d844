#!/bin/bash
# same-box A/B of two builds of libfsb.so: build/wt_prev_so/libfsb.so (copied before a change) against the in-tree one.
# First the wide-batch parity tests on the new build, then cfg5 / cfg3 decode with the kernel's sub-timers.
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_lm_gpu.py -x -q -m gpu -k "wide or batch or session or snapshot or ragged" > gpurun_out/ab_pytest.log 2>&1
tail -3 gpurun_out/ab_pytest.log
for rep in 1 2; do
for v in cur prev; do
  if [ $v = prev ]; then export FSB_LIB=$PWD/build/wt_prev_so/libfsb.so; else unset FSB_LIB; fi
  for cfg in cfg5 cfg3; do
  timeout -s KILL 300 python bench.py --config $cfg --steps 2 --warmup 1 --no-cpu-baseline --no-extras 2>gpurun_out/ab_${v}_${cfg}.err | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$v rep$rep $cfg', round(d['value']), {k:round(x,1) for k,x in d['breakdown_ms_per_step'].items()})"
  done
done
done
unset FSB_LIB
FSB_MEGA_TIMERS=1 timeout -s KILL 300 python bench.py --config cfg5 --steps 1 --warmup 1 --no-cpu-baseline --no-extras 2>&1 >/dev/null | grep "\[megab\|block_sample" | tail -12

#!/bin/bash
# same-box A/B: the library of an earlier commit (build/wt_prev) against the current one
mkdir -p gpurun_out
for rep in 1 2; do
for v in cur prev; do
  if [ $v = prev ]; then D=build/wt_prev; else D=.; fi
  (cd $D && timeout -s KILL 400 python bench.py --config cfg5 --steps 2 --warmup 1 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$v rep$rep cfg5', round(d['value']), {k:round(x) for k,x in d['breakdown_ms_per_step'].items()})")
done
done

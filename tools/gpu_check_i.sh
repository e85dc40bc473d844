#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_lm_gpu.py tests/test_ops_gpu.py -m gpu -x -q --timeout 100 2>&1 | tail -4
run() { timeout -s KILL 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$1', d['value'], d['breakdown_ms_per_step'])"; }
run new

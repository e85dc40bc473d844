#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 240 python -m pytest tests/test_lm_gpu.py -m gpu -x -q --timeout 100 -k "single_row_ring" 2>&1 | tail -3
run() { timeout -s KILL 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$1', d['value'], d['breakdown_ms_per_step'])"; }
run ll
FSB_MEGA_BARRIER=1 run barrier
FSB_MEGA_TIMERS=1 timeout -s KILL 200 python bench.py --steps 1 --warmup 1 --frames 32 --no-cpu-baseline 2>&1 >/dev/null | grep -E "mega cta 0" | tail -7

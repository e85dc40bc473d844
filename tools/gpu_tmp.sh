#!/bin/bash
timeout -s KILL 300 python -m pytest tests/test_lm_gpu.py -m gpu -x -q --timeout 200 -k "single_row_ring or full_size" 2>&1 | tail -2
for i in 1 2; do timeout -s KILL 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('spec', d['value'], d['breakdown_ms_per_step'])"; done

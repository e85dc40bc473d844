#!/bin/bash
mkdir -p gpurun_out
run() { timeout -s KILL 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$1', d['value'], d['breakdown_ms_per_step'])"; }
run cta0
FSB_MEGA_SAMPLER_CTA=1 run last
FSB_MEGA_V1=1 run v1

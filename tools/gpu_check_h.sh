#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_codec_gpu.py -m gpu -q --timeout 120 2>&1 | tail -2
run() { timeout -s KILL 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$1', d['value'], d['e2e']['value'], d['breakdown_ms_per_step'])"; }
run wide
FSB_CONV_NARROW=1 run narrow

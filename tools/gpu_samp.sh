#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests/test_lm_gpu.py -m gpu -q --timeout 600 > gpurun_out/pytest_lm.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_lm.log
FSB_MEGA_TIMERS=1 timeout -s KILL 400 python bench.py --config cfg5 --frames 200 --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/cfg5_timers_s.json 2> gpurun_out/cfg5_timers_s.err; grep "sample_fast\|block_sample\|cta 0\] sample" gpurun_out/cfg5_timers_s.err
timeout -s KILL 400 python bench.py --config cfg2 --steps 3 --warmup 2 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('cfg2', d['value'],d['breakdown_ms_per_step'],d['roofline']['frac'])"
timeout -s KILL 400 python bench.py --config cfg5 --steps 2 --warmup 1 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('cfg5', d['value'],d['breakdown_ms_per_step'],d['roofline']['frac'])"

#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_lm_gpu.py -m gpu -q --timeout 600 -k "wide or ragged or session or snapshot or full_size" > gpurun_out/pytest_megab.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_megab.log
FSB_MEGA_TIMERS=1 timeout -s KILL 400 python bench.py --config cfg3 --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/cfg3_timers_0.json 2> gpurun_out/cfg3_timers_0.err; grep "cta 0\] w2\|cta 74\] w2" gpurun_out/cfg3_timers_0.err
timeout -s KILL 400 python bench.py --config cfg3 --steps 3 --warmup 1 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('cfg3', d['value'],d['breakdown_ms_per_step'],d['roofline']['frac'])"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_cfg3_prefill.csv python bench.py --config cfg3 --frames 2 --steps 1 --warmup 0 --no-cpu-baseline --no-extras > /dev/null 2>&1; wc -l gpurun_out/launches_cfg3_prefill.csv

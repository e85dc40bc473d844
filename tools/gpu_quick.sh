#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_lm_gpu.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_lm.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_lm.log
timeout -s KILL 300 python bench.py --config cfg3 --steps 2 --warmup 1 --no-cpu-baseline --no-extras 2>gpurun_out/q3.err | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('cfg3', d['value'], d['breakdown_ms_per_step'])"
timeout -s KILL 300 python bench.py --config cfg5 --frames 64 --steps 1 --warmup 1 --no-cpu-baseline --no-extras 2>gpurun_out/q5.err | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('cfg5/64', d['value'], d['breakdown_ms_per_step'])"

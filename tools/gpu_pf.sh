#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests/test_lm_gpu.py -m gpu -q --timeout 600 -x > gpurun_out/pytest_lm.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_lm.log
timeout -s KILL 400 python bench.py --config cfg3 --steps 3 --warmup 1 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('cfg3', d['value'],d['breakdown_ms_per_step'],d['roofline']['frac'])"
FSB_PREFILL_ATT_V1=1 timeout -s KILL 400 python bench.py --config cfg3 --steps 3 --warmup 1 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('cfg3 v1', d['value'],d['breakdown_ms_per_step'],d['roofline']['frac'])"
timeout -s KILL 400 python bench.py --config cfg5 --steps 2 --warmup 1 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('cfg5', d['value'],d['breakdown_ms_per_step'],d['roofline']['frac'])"
timeout -s KILL 400 python bench.py --config cfg2 --steps 3 --warmup 2 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('cfg2', d['value'],d['breakdown_ms_per_step'],d['roofline']['frac'])"

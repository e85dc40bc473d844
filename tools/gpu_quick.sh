#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_lm_gpu.py -m gpu -q -x --timeout 600 -k "session or snapshot or wide_batch or natural_stop" > gpurun_out/pytest_lm.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_lm.log

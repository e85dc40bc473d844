#!/bin/bash
# full GPU evidence run: parity tests, smoke, default bench (cfg5 shard + secondary blocks), reference arm
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log

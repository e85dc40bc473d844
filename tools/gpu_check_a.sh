#!/bin/bash
# Re-entry validation: parity tests, bench cfg2 bf16, phase timers, cfg3 batch bench.
set -x
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/pytest_gpu_a.log 2>&1; tail -5 gpurun_out/pytest_gpu_a.log
timeout -s KILL 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; cat gpurun_out/bench_a.json
FSB_MEGA_TIMERS=1 timeout -s KILL 200 python bench.py --steps 1 --warmup 1 --frames 32 --no-cpu-baseline > gpurun_out/bench_timers.json 2> gpurun_out/bench_timers.err; tail -40 gpurun_out/bench_timers.err
timeout -s KILL 300 python bench.py --config cfg3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_cfg3_a.json 2> gpurun_out/bench_cfg3_a.err; cat gpurun_out/bench_cfg3_a.json

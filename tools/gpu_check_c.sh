#!/bin/bash
set -x
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_lm_gpu.py -m gpu -x -q --timeout 100 -k "single_row_ring or decode_modes or bf16_weights" > gpurun_out/pytest_c.log 2>&1; tail -25 gpurun_out/pytest_c.log
timeout -s KILL 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; cat gpurun_out/bench_c.json; tail -3 gpurun_out/bench_c.err
FSB_MEGA_TIMERS=1 timeout -s KILL 200 python bench.py --steps 1 --warmup 1 --frames 32 --no-cpu-baseline > gpurun_out/bench_timers_c.json 2> gpurun_out/bench_timers_c.err; tail -22 gpurun_out/bench_timers_c.err

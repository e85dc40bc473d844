#!/bin/bash
set -x
mkdir -p gpurun_out
FSB_MEGA_TIMERS=1 timeout -s KILL 200 python bench.py --steps 1 --warmup 1 --frames 32 --no-cpu-baseline > gpurun_out/bench_timers_d.json 2> gpurun_out/bench_timers_d.err; tail -22 gpurun_out/bench_timers_d.err

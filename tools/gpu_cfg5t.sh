#!/bin/bash
mkdir -p gpurun_out
FSB_MEGA_TIMERS=1 timeout -s KILL 400 python bench.py --config cfg5 --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/cfg5_timers.json 2> gpurun_out/cfg5_timers.err; grep megab gpurun_out/cfg5_timers.err | tail -5; grep "mega cta" gpurun_out/cfg5_timers.err | tail -14

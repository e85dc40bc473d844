#!/bin/bash
mkdir -p gpurun_out
FSB_MEGA_TIMERS=1 timeout -s KILL 300 python bench.py --config cfg3 --steps 1 --warmup 1 --no-cpu-baseline --frames 32 > gpurun_out/cfg3_timers.json 2> gpurun_out/cfg3_timers.err; grep megab gpurun_out/cfg3_timers.err | tail -8; grep "mega cta 0" gpurun_out/cfg3_timers.err | tail -7

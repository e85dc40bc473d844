#!/bin/bash
mkdir -p gpurun_out
FSB_MEGA_TIMERS=1 timeout -s KILL 300 python bench.py --config cfg3 --steps 1 --warmup 1 --no-cpu-baseline --frames 32 > gpurun_out/cfg3_timers.json 2> gpurun_out/cfg3_timers.err; tail -30 gpurun_out/cfg3_timers.err
timeout -s KILL 400 python bench.py --config cfg3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/cfg3_r2a.json 2> gpurun_out/cfg3_r2a.err; tail -3 gpurun_out/cfg3_r2a.err; cat gpurun_out/cfg3_r2a.json

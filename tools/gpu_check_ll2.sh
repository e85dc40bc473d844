#!/bin/bash
mkdir -p gpurun_out
FSB_MEGA_TIMERS=1 timeout -s KILL 200 python bench.py --steps 1 --warmup 1 --frames 32 --no-cpu-baseline 2>&1 >/dev/null | grep -E "mega cta" | tail -14

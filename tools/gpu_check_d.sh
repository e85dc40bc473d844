#!/bin/bash
FSB_MEGA_TIMERS=1 timeout -s KILL 60 python bench.py --steps 1 --warmup 1 --frames 32 --no-cpu-baseline 2>&1 >/dev/null | grep -E "sampler dbg|block_sample|sample " | tail -5

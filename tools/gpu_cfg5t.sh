#!/bin/bash
mkdir -p gpurun_out
FSB_MEGA_TIMERS=1 timeout -s KILL 400 python bench.py --config cfg5 --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/cfg5_timers.json 2> gpurun_out/cfg5_timers.err; grep "megab owner" gpurun_out/cfg5_timers.err | tail -4; tail -2 gpurun_out/cfg5_timers.err; python -c "
import json;d=json.load(open('gpurun_out/cfg5_timers.json'));print(d['value'],d['breakdown_ms_per_step'],d['roofline']['frac'])"
FSB_MEGA_TIMERS=1 timeout -s KILL 400 python bench.py --config cfg3 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/cfg3_timers.json 2> gpurun_out/cfg3_timers.err; grep "megab owner" gpurun_out/cfg3_timers.err | tail -4; python -c "
import json;d=json.load(open('gpurun_out/cfg3_timers.json'));print(d['value'],d['breakdown_ms_per_step'],d['roofline']['frac'])"

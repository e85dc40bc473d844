#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_cfg5_prefill_r2.csv python bench.py --config cfg5 --frames 2 --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/ncu_launches.log 2>&1; wc -l gpurun_out/launches_cfg5_prefill_r2.csv

#!/bin/bash
set -x
mkdir -p gpurun_out
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k regex:mega1_decode -c 1 -f -o gpurun_out/mega1_r1 python bench.py --steps 1 --warmup 0 --frames 8 --no-cpu-baseline > gpurun_out/ncu_mega1.log 2>&1
tail -3 gpurun_out/ncu_mega1.log
ls -la gpurun_out/*.ncu-rep

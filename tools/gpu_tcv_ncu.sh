#!/bin/bash
mkdir -p gpurun_out
T=1292 REPS=1 timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_codec_tc_1292.csv python tools/gpu_codec_profile.py > /dev/null 2>&1; wc -l gpurun_out/launches_codec_tc_1292.csv

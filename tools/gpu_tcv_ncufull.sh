#!/bin/bash
mkdir -p gpurun_out
# one full capture of a stage-1 ResBlock conv (C = 128, K = 11) and one of a stage-3 conv (C = 32) at T = 1292
T=1292 REPS=1 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:tcconv_kernel -s 36 -c 1 -f -o gpurun_out/tcconv128_r2 python tools/gpu_codec_profile.py > gpurun_out/ncu_tcconv128.log 2>&1; tail -2 gpurun_out/ncu_tcconv128.log
T=1292 REPS=1 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:tcconv_kernel -s 75 -c 1 -f -o gpurun_out/tcconv32_r2 python tools/gpu_codec_profile.py > gpurun_out/ncu_tcconv32.log 2>&1; tail -2 gpurun_out/ncu_tcconv32.log
T=1292 REPS=1 timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_codec_tc_1292.csv python tools/gpu_codec_profile.py > /dev/null 2>&1; wc -l gpurun_out/launches_codec_tc_1292.csv
T=216 REPS=1 timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_codec_tc_216.csv python tools/gpu_codec_profile.py > /dev/null 2>&1; wc -l gpurun_out/launches_codec_tc_216.csv

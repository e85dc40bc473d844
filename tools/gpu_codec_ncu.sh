#!/bin/bash
mkdir -p gpurun_out
T=216 python tools/gpu_codec_profile.py | tail -1
T=1292 python tools/gpu_codec_profile.py | tail -1
T=216 REPS=1 timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:resconv_tma_kernel -s 34 -c 1 -f -o gpurun_out/resconv_r2 python tools/gpu_codec_profile.py > gpurun_out/ncu_resconv.log 2>&1; tail -2 gpurun_out/ncu_resconv.log
T=216 REPS=1 timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_codec_r2.csv python tools/gpu_codec_profile.py > /dev/null 2>&1; wc -l gpurun_out/launches_codec_r2.csv

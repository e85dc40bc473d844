#!/bin/bash
# final evidence of round 2 (session 2): parity tests, smoke, reference arm, default bench, then the ncu capture of the
# wide-batch decode kernel (one launch, 8 frames) and the launch list of the same command
bash tools/gpu_final.sh
B="python bench.py --config cfg5 --frames 8 --steps 1 --warmup 0 --no-cpu-baseline --no-extras"
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:megab_decode -c 1 -f -o gpurun_out/megab_r2e $B > gpurun_out/ncu_megab_r2e.log 2>&1; tail -2 gpurun_out/ncu_megab_r2e.log
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_cfg5_frames8_r2e.csv $B > gpurun_out/ncu_launches_r2e.log 2>&1; wc -l gpurun_out/launches_cfg5_frames8_r2e.csv

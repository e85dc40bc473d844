#!/bin/bash
mkdir -p gpurun_out
ROWS=16 FRAMES=24 MODES=2 timeout -s KILL 300 python tools/gpu_megab_bringup.py > gpurun_out/megab_bringup2.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/megab_bringup2.log
ROWS=27 FRAMES=12 MODES=2 timeout -s KILL 300 python tools/gpu_megab_bringup.py > gpurun_out/megab_bringup3.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/megab_bringup3.log
bash tools/gpu_cfg5t.sh

#!/bin/bash
# bring-up of the wide-batch megakernel
mkdir -p gpurun_out
ROWS=11 FRAMES=6 timeout -s KILL 200 python tools/gpu_megab_bringup.py > gpurun_out/megab_bringup.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/megab_bringup.log
ROWS=16 FRAMES=48 timeout -s KILL 400 python tools/gpu_megab_bringup.py > gpurun_out/megab_bringup2.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/megab_bringup2.log
ROWS=27 FRAMES=20 MODES=2 timeout -s KILL 400 python tools/gpu_megab_bringup.py > gpurun_out/megab_bringup3.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/megab_bringup3.log

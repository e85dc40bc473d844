#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python tools/gpu_tcv_bringup.py > gpurun_out/tcv_bringup.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/tcv_bringup.log

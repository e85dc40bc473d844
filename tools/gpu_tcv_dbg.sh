#!/bin/bash
mkdir -p gpurun_out
FSB_TCV_DBG=1 TS=1292 timeout -s KILL 300 python tools/gpu_tcv_bringup.py > gpurun_out/tcv_dbg.log 2>&1; echo "rc=$?"; grep "tcconv N" gpurun_out/tcv_dbg.log | head -100 > gpurun_out/tcv_dbg_first.log

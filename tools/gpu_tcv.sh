#!/bin/bash
mkdir -p gpurun_out
TS=8,216,1292 timeout -s KILL 300 python tools/gpu_tcv_bringup.py > gpurun_out/tcv_bringup.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/tcv_bringup.log
timeout -s KILL 900 python -m pytest tests/test_codec_gpu.py -m gpu -q --timeout 300 > gpurun_out/pytest_codec.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_codec.log

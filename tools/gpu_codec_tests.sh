#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_codec_gpu.py -m gpu -q --timeout 300 > gpurun_out/pytest_codec.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_codec.log

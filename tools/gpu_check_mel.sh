#!/bin/bash
timeout -s KILL 300 python -m pytest tests/test_codec_gpu.py tests/test_abi.py -m gpu -q --timeout 120 2>&1 | tail -15

#!/bin/bash
timeout -s KILL 500 python -m pytest tests/test_lm_gpu.py -m gpu -q --timeout 300 -k "cfg4_voice_clone or full_size_properties" 2>&1 | tail -25

#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_codec_gpu.py -m gpu -q -x -k "decode_matches_oracle or tensor_core or block_decode" > gpurun_out/sanitizer_codec.log 2>&1; echo "memcheck rc=$?"; grep -c "Invalid\|out of bounds" gpurun_out/sanitizer_codec.log; tail -4 gpurun_out/sanitizer_codec.log
timeout -s KILL 300 python bench.py --config cfg2 --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/b1_stdout.txt 2> gpurun_out/b1_stderr.txt; echo "stdout lines: $(wc -l < gpurun_out/b1_stdout.txt)"; head -c 120 gpurun_out/b1_stdout.txt; echo

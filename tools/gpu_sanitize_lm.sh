#!/bin/bash
# compute-sanitizer memcheck of the wide-batch decode path (NPAD 16 and 32) and the batched prefill in front of it
mkdir -p gpurun_out
timeout -s KILL 240 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_lm_gpu.py -m gpu -q -x -k "wide_batch_megakernel_matches_oracle and (11-8 or 27-12)" > gpurun_out/sanitizer_lm.log 2>&1; echo "memcheck rc=$?"; grep -c "Invalid\|out of bounds" gpurun_out/sanitizer_lm.log; tail -4 gpurun_out/sanitizer_lm.log

#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 170 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 python -m pytest tests/test_lm_gpu.py -m gpu -q -x -k "wide_batch_megakernel_matches_oracle and 11-8" > gpurun_out/racecheck_lm.log 2>&1; echo "racecheck rc=$?"; grep -c "hazard" gpurun_out/racecheck_lm.log; grep "hazard" gpurun_out/racecheck_lm.log | sort | uniq -c | sort -rn | head -8; tail -3 gpurun_out/racecheck_lm.log

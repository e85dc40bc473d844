#!/bin/bash
# round-2 ncu evidence: wide-batch decode kernel (full set), tcgen05 prefill GEMM (full set), launch list of one cfg5 step (8 frames)
mkdir -p gpurun_out
B="python bench.py --config cfg5 --frames 8 --steps 1 --warmup 0 --no-cpu-baseline --no-extras"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:megab_decode -c 1 -f -o gpurun_out/megab_r2 $B > gpurun_out/ncu_megab.log 2>&1; tail -2 gpurun_out/ncu_megab.log
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 40 -c 1 -f -o gpurun_out/tcgemm_r2 $B > gpurun_out/ncu_tcgemm.log 2>&1; tail -2 gpurun_out/ncu_tcgemm.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_cfg5_frames8_r2.csv $B > gpurun_out/ncu_launches.log 2>&1; wc -l gpurun_out/launches_cfg5_frames8_r2.csv

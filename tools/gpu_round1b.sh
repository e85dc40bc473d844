#!/bin/bash
# Round-1 (second half) GPU evidence: parity tests, smoke, bench (ours + reference arm + f32), launch list, ncu captures.
set -x
mkdir -p gpurun_out
timeout -s KILL 500 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout -s KILL 500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; cat gpurun_out/bench_r1b.json
timeout -s KILL 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_r1b.json 2> gpurun_out/bench_ref_r1b.err; cat gpurun_out/bench_ref_r1b.json
timeout -s KILL 300 python bench.py --steps 3 --warmup 3 --dtype f32 --no-cpu-baseline > gpurun_out/bench_f32_r1b.json 2>/dev/null; cat gpurun_out/bench_f32_r1b.json
timeout -s KILL 300 python bench.py --config cfg3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_cfg3_r1b.json 2>/dev/null; cat gpurun_out/bench_cfg3_r1b.json
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 1 --warmup 0 --frames 8 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:mega1_decode -c 1 -f -o gpurun_out/mega1_r1b python bench.py --steps 1 --warmup 0 --frames 8 --no-cpu-baseline > gpurun_out/ncu_mega1.log 2>&1
# conv: the largest ResBlock conv of a full-length (216 frame) vocoder pass (stage 1, k = 11: launch 58 of the codec)
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:conv1d_kernel -s 58 -c 1 -f -o gpurun_out/conv_r1b python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_conv.log 2>&1
ls -la gpurun_out

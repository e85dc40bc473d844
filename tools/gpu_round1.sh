#!/bin/bash
# Round-1 GPU evidence run: parity tests, bench (ours + reference arm), ncu launch list and full captures.
set -x
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout -s KILL 500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; cat gpurun_out/bench_r1.json
timeout -s KILL 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_r1.json 2> gpurun_out/bench_ref_r1.err; cat gpurun_out/bench_ref_r1.json
timeout -s KILL 300 python bench.py --steps 3 --warmup 3 --dtype f32 --no-cpu-baseline > gpurun_out/bench_f32_r1.json 2>/dev/null; cat gpurun_out/bench_f32_r1.json
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 0 --frames 8 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:mega_decode -c 1 -o gpurun_out/mega_r1 python bench.py --steps 1 --warmup 0 --frames 8 --no-cpu-baseline > gpurun_out/ncu_mega.log 2>&1
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:conv1d_kernel -s 40 -c 3 -o gpurun_out/conv_r1 python bench.py --steps 1 --warmup 0 --frames 8 --no-cpu-baseline > gpurun_out/ncu_conv.log 2>&1
ls -la gpurun_out

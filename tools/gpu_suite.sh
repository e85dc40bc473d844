#!/bin/bash
# full GPU evidence run: parity tests, smoke, default bench (cfg5 shard + secondary blocks)
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout -s KILL 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; tail -3 gpurun_out/bench_r2c.err; python -c "
import json;d=json.load(open('gpurun_out/bench_r2c.json'));print(d['value'], d['e2e']['value'], d['breakdown_ms_per_step'], 'frac', d['roofline']['frac'], d['roofline']['frame_level_frac'], 'voc', d['vocoder_roofline']['achieved']); print('cfg3', d['cfg3']['value'], d['cfg3']['breakdown_ms_per_step']); print('b1', d['latency_b1']['value'], d['latency_b1']['breakdown_ms_per_step']); print(d['cpu_baseline'])"

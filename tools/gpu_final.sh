#!/bin/bash
# final evidence run of the round: parity tests, smoke, reference arm, default bench (cfg5 shard + secondary blocks)
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout -s KILL 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; python -c "
import json;d=json.load(open('gpurun_out/bench_final_ref.json'));print('reference arm', d['value'], d['cpu_baseline']['cores'], 'threads')"
timeout -s KILL 900 python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; tail -2 gpurun_out/bench_final_n1.err; python -c "
import json;d=json.load(open('gpurun_out/bench_final_n1.json'));print(d['value'], d['e2e']['value'], d['breakdown_ms_per_step'], 'frac', d['roofline']['frac'], 'voc', d['vocoder_roofline']['achieved'], d['vocoder_roofline']['frac'], 'launches', d['gpu_launches']); print('cfg3', d['cfg3']['value'], d['cfg3']['breakdown_ms_per_step']); print('b1', d['latency_b1']['value'], d['latency_b1']['breakdown_ms_per_step']); print(d['cpu_baseline']); print(d['clocks'])"

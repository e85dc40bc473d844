#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_codec_gpu.py tests/test_lm_gpu.py -m gpu -q -x --timeout 300 -k "codec or hidden or cfg4 or decode or block or sky or fsq" > gpurun_out/pytest_codec.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -6 gpurun_out/pytest_codec.log
if [ $rc != 0 ]; then
  timeout -s KILL 300 compute-sanitizer --tool memcheck python -m pytest tests/test_codec_gpu.py -m gpu -q -x -k "decode_matches_oracle and 1" > gpurun_out/sanitizer.log 2>&1; grep -v "^$" gpurun_out/sanitizer.log | head -40
  exit 0
fi
for v in 0 1; do
  if [ $v = 1 ]; then export FSB_CODEC_NO_TMA=1; fi
  timeout -s KILL 300 python bench.py --config cfg3 --steps 2 --warmup 1 --no-cpu-baseline --no-extras 2>gpurun_out/codec_bench_$v.err | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('NO_TMA=$v', d['value'], d['breakdown_ms_per_step'], d['vocoder_roofline']['achieved'])"
done

#!/bin/bash
timeout -s KILL 300 python -m pytest tests/test_codec_gpu.py tests/test_abi.py -m gpu -q --timeout 120 2>&1 | tail -15
timeout -s KILL 120 python - <<'PY'
import numpy as np, time
from fish_speech_rs_b200 import FireflyCodec, synth
from oracle import mel as omel
w = synth.make_codec_weights(seed=4321, with_encoder=True)
c = FireflyCodec(w, max_frames=300, with_encoder=True)
rng = np.random.default_rng(0)
n = 562265
pcm = (0.1 * rng.standard_normal(n)).astype(np.float32)
for _ in range(2):
    t = time.perf_counter(); m = c.log_mel(pcm); t1 = time.perf_counter() - t
t = time.perf_counter(); e = omel.log_mel(pcm); t2 = time.perf_counter() - t
print("sky-sized clip: frames", m.shape, "gpu %.2f ms, oracle %.1f ms, max abs diff %.2e" % (t1 * 1e3, t2 * 1e3, np.abs(m[0] - e).max()))
t = time.perf_counter(); codes = c.encode(pcm); print("encode", codes.shape, "%.2f ms" % ((time.perf_counter() - t) * 1e3))
PY

"""Vocoder-only run for profiling: one utterance of T frames through FireflyCodec.decode."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fish_speech_rs_b200 import FireflyCodec, synth
T = int(os.environ.get("T", "216"))
w = synth.make_codec_weights(seed=4321, with_encoder=False)
codec = FireflyCodec(w, max_frames=T)
codes = np.random.default_rng(7).integers(0, 1000, size=(1, 8, T)).astype(np.uint32)
for i in range(int(os.environ.get("REPS", "3"))):
    t0 = time.time()
    codec.decode(codes)
    st = codec.stats()
    print(f"T={T} device_ms {st['device_ms']:.2f} launches {st['kernel_launches']} -> {2.646e9 * T / st['device_ms'] / 1e9:.1f} TFLOP/s")

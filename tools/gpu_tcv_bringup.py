"""Bring-up of the tcgen05 ResBlock convs: PCM of the tensor-core path vs the FP32-FMA path vs the oracle, and timings."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fish_speech_rs_b200 import FireflyCodec, synth

w = synth.make_codec_weights(seed=4321, with_encoder=False)
Ts = [int(x) for x in os.environ.get("TS", "8,216,1292").split(",")]


def run(T, tc, reps=3):
    if tc:
        os.environ.pop("FSB_CODEC_NO_TC", None)
    else:
        os.environ["FSB_CODEC_NO_TC"] = "1"
    codec = FireflyCodec(w, max_frames=T)
    codes = np.random.default_rng(7).integers(0, 1000, size=(1, 8, T)).astype(np.uint32)
    ms = []
    for _ in range(reps):
        pcm = codec.decode(codes)
        ms.append(codec.stats()["device_ms"])
    n = codec.stats()["kernel_launches"]
    codec.close()
    return pcm, min(ms), n, codes


for T in Ts:
    a, ms_a, n_a, codes = run(T, True)
    b, ms_b, n_b, _ = run(T, False)
    print(f"T={T}: tc {ms_a:.2f} ms ({2.646 * T / ms_a:.1f} TFLOP/s) vs fma {ms_b:.2f} ms ({2.646 * T / ms_b:.1f} TFLOP/s); "
          f"max|tc - fma| = {np.abs(a - b).max():.3e} (max|pcm| {np.abs(b).max():.3f}), nan={np.isnan(a).any()}", flush=True)
    if T <= 16:
        import torch
        from oracle import codec as ocodec
        with torch.no_grad():
            ref = ocodec.decode(torch.from_numpy(codes.astype(np.int64)), w).numpy()
        print(f"   vs oracle: tc {np.abs(a - ref).max():.3e}  fma {np.abs(b - ref).max():.3e}", flush=True)

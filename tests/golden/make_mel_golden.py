#!/usr/bin/env python3
"""Pins oracle/mel.py against the reference's own data.  Runs only where /root/reference is mounted (the build
container); writes tests/golden/mel_golden.json, which the CPU tests read instead of the reference tree.
  * audio/melfilters160.bytes (1025 x 160 little-endian f32): max-abs difference to oracle.mel.mel_filterbank()
  * tests/resources/sky.wav (562 265 samples): mel frame count (1099) and the code-frame arithmetic (274)
  * a digest of the oracle's filterbank so that later edits of the formula are caught without the reference tree"""
import hashlib
import json
import os
import sys
import wave

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import mel  # noqa: E402

REF = "/root/reference"
ref_fb = np.fromfile(os.path.join(REF, "fish_speech_core/lib/audio/melfilters160.bytes"), dtype="<f4").reshape(1025, 160)
fb = mel.mel_filterbank()
with wave.open(os.path.join(REF, "tests/resources/sky.wav")) as w:
    n = w.getnframes()
    sr = w.getframerate()
lm = mel.n_mel_frames(n)
l1 = (lm - 2) // 2 + 1
out = {
    "melfilters160_max_abs_diff": float(np.abs(ref_fb - fb).max()),
    "melfilters160_ref_sha256": hashlib.sha256(ref_fb.tobytes()).hexdigest(),
    "oracle_fb_sha256": hashlib.sha256(fb.tobytes()).hexdigest(),
    "oracle_fb_col_sums_first8": [float(v) for v in fb.sum(0)[:8]],
    "oracle_fb_nonzeros": int((fb > 0).sum()),
    "sky_wav": {"samples": n, "sample_rate": sr, "mel_frames": lm, "code_frames": (l1 - 2) // 2 + 1},
}
json.dump(out, open(os.path.join(ROOT, "tests/golden/mel_golden.json"), "w"), indent=1)
print(json.dumps(out, indent=1))

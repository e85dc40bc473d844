#!/usr/bin/env python3
"""Copies the loader ground truth the reference holds into tests/golden/ so that the CPU tests can pin
`synth.lm_weight_shapes` / `synth.codec_weight_shapes` (and with them the loaders of csrc/) without the reference tree:
  docs/llama-weight-dict.txt    -> llama_weight_dict_fish12.txt   (LM tensor names + shapes, Fish 1.2 checkpoint)
  docs/weight-dims-default.txt  -> codec_weight_dims_fish12.txt   (codec names + shapes, Fish 1.2, un-merged weight norm)
  tests/resources/sky.wav       -> sky.wav                         (12.75 s mono 44.1 kHz s16: the encoder's known input)
Runs only where /root/reference is mounted (the build container)."""
import hashlib
import os
import shutil

ROOT = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
for src, dst in (("docs/llama-weight-dict.txt", "llama_weight_dict_fish12.txt"),
                 ("docs/weight-dims-default.txt", "codec_weight_dims_fish12.txt"),
                 ("tests/resources/sky.wav", "sky.wav")):
    shutil.copyfile(os.path.join(REF, src), os.path.join(ROOT, dst))
    print(dst, hashlib.sha256(open(os.path.join(ROOT, dst), "rb").read()).hexdigest())

import sys, numpy as np
sys.path.insert(0, '.')
from fish_speech_rs_b200 import DualARTransformer, SamplingArgs, generate_blocking, generate_blocking_with_hidden, synth
cfg, tok = dict(synth.TINY), dict(synth.TINY_TOKENS)
w = synth.make_lm_weights(cfg, seed=1234)
gpu = DualARTransformer(w, cfg, tok, max_batch=4, max_seq_len=512)
prompt = synth.make_prompt(cfg, tok, 26, seed=314)
lg, hg = gpu.forward_generate(prompt[None], 0); gpu.clear_slow_layer_caches()
try:
    c, h = generate_blocking_with_hidden(gpu, prompt, 400, SamplingArgs(temp=0.0), fixed_len=9)
    print(c.shape, h.shape)
except Exception as e:
    print("ERR", e)

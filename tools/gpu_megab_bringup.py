"""Bring-up check of the wide-batch tcgen05 megakernel (fsb_lm_megab.cuh): every sampling decision of every row is
replayed against the oracle under teacher forcing (oracle/generate.py::replay_frames)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fish_speech_rs_b200 import DualARTransformer, SamplingArgs, generate_static_batch, synth
from oracle import dual_ar as olm, generate as ogen, sampling as osamp

def t64(a):
    return torch.from_numpy(np.asarray(a).astype(np.int64))

cfg, tok = dict(synth.WIDE), dict(synth.TINY_TOKENS)
w = synth.make_lm_weights(cfg, seed=77, round_bf16=True)
nrows = int(os.environ.get("ROWS", "11"))
nfr = int(os.environ.get("FRAMES", "6"))
modes = [int(m) for m in os.environ.get("MODES", "2,1").split(",")]
prompts = [synth.make_prompt(cfg, tok, 12 + 7 * i, seed=300 + i) for i in range(nrows)]
ora = olm.DualARTransformer(w, olm.BaseModelArgs(**cfg), olm.TokenConfig(**tok), "1.5")
worst = 0
for mode in modes:
    gpu = DualARTransformer(w, cfg, tok, max_batch=nrows, max_seq_len=256, decode_mode=mode, dtype="bf16")
    for name, sa, so in (("greedy", SamplingArgs(temp=0.0), osamp.SamplingArgs(temp=0.0)),
                         ("sampled", SamplingArgs(0.7, 0.8, 256, 1.4, seed=9), osamp.SamplingArgs(0.7, 0.8, 256, 1.4, seed=9))):
        t0 = time.time()
        out = generate_static_batch(gpu, prompts, 400, sa, fixed_len=nfr)
        st = gpu.stats()
        tot = dict(decisions=0, exact=0, near_tie=0, violation=0)
        for i in range(nrows):
            fr = gpu.last_frames(i)
            assert fr.shape[1] == nfr and np.array_equal(fr[1:], out[i]), (fr.shape, out[i].shape)
            r = ogen.replay_frames(ora, t64(prompts[i]), fr, so, row=i, fixed_len=nfr)
            for k in tot:
                tot[k] += r[k]
            if r["violation"]:
                print(f"  mode {mode} {name} row {i}: violations at {r['events'][:6]}")
        worst = max(worst, tot["violation"])
        print(f"mode {mode} {name}: decode {st['decode_ms']:.2f} ms prefill {st['prefill_ms']:.2f} ms launches {st['kernel_launches']} "
              f"-> {tot}  ({time.time() - t0:.1f}s)", flush=True)
    gpu.close()
sys.exit(1 if worst else 0)

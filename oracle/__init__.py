"""CPU restatement ("oracle") of the fish-speech.rs hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``fish_speech_rs_b200/`` imports this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may.

PARITY UNPINNED: the reference (Rust + Candle 0.8.3) can be neither compiled
nor imported in this container (no cargo/rustc, no network, no weights) and its
own tests hold no golden vectors for this path.  The
oracle is therefore a PyTorch CPU fp32 restatement written op-for-op from

  fish_speech_core/lib/lm/dual_ar.rs
  fish_speech_core/lib/lm/generate/{single_batch,static_batch,utils}.rs
  fish_speech_core/lib/lm/sampling/{mod,rep_pen}.rs
  fish_speech_core/lib/codec/*.rs
  fish_speech_core/lib/audio/{spectrogram,stft}.rs

pinned by what the reference does carry -- its loader dumps
(docs/llama-weight-dict.txt, docs/weight-dims-default.txt), its mel filterbank
bytes, tests/resources/sky.wav, voices-template/default.npy -- and by the
closed forms in its source (FSQ implicit codebook, causal-mask truth table, id
rescaling round trip, rep-pen window behaviour, repeat_kv test shapes): see
``tests/test_oracle_kat.py`` and DESIGN.md section 5.  The float arithmetic of
the two transformer stacks is additionally tied to an independently written
Llama (Hugging Face ``transformers``, ``tests/test_oracle_vs_hf_llama.py``:
2e-6 after mapping the interleaved RoPE pairs); that is a check of the
restatement, not of the reference, so the label above stays.
"""

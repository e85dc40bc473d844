"""GPU parity tests of the dual-AR token loop against the oracle, through the C ABI.
Tolerances: logits / hidden atol 1e-3 (the reference author's own bar, tests/e2e/
backbone-allclose.py:82); token ids bit-exact under greedy decode (north_star)."""
import os

import numpy as np
import pytest
import torch

from fish_speech_rs_b200 import DualARTransformer, SamplingArgs, generate_blocking, generate_static_batch, synth
from oracle import dual_ar as olm
from oracle import generate as ogen
from oracle import sampling as osamp

pytestmark = pytest.mark.gpu
ATOL = 1e-3


def oracle_model(cfg, tok, w, version="1.5"):
    return olm.DualARTransformer(w, olm.BaseModelArgs(**cfg), olm.TokenConfig(**tok), version)


def t64(a):
    return torch.from_numpy(np.asarray(a).astype(np.int64))


@pytest.fixture(scope="module")
def models(tiny_lm):
    cfg, tok, w = tiny_lm
    gpu = DualARTransformer(w, cfg, tok, max_batch=4, max_seq_len=512)
    yield cfg, tok, w, gpu, oracle_model(cfg, tok, w)
    gpu.close()


def test_prefill_and_decode_steps_match_oracle(models):
    cfg, tok, w, gpu, ora = models
    C = cfg["num_codebooks"]
    for P in (1, 7, 33, 166):  # 166 mirrors repeat_kv.rs:148
        gpu.clear_slow_layer_caches()
        ora.clear_slow_layer_caches()
        prompt = synth.make_prompt(cfg, tok, max(P, 12), seed=5)[:, -P:]
        with torch.no_grad():
            lo, ho = ora.forward_generate(t64(prompt)[None], 0)
        lg, hg = gpu.forward_generate(prompt[None], 0)
        np.testing.assert_allclose(hg, ho.numpy(), atol=ATOL, rtol=0)
        np.testing.assert_allclose(lg, lo.numpy(), atol=ATOL, rtol=0)
        assert gpu.curr_kv_size() == ora.curr_kv_size() == P
        # three decode steps on top of the prefilled KV
        rng = np.random.default_rng(P)
        for s in range(3):
            step = np.zeros((1, C + 1, 1), np.uint32)
            step[0, 0, 0] = tok["semantic_start_id"] + rng.integers(0, 1024)
            step[0, 1:, 0] = rng.integers(0, 1024, size=C)
            with torch.no_grad():
                lo, ho = ora.forward_generate(t64(step), P + s)
            lg, hg = gpu.forward_generate(step, P + s)
            np.testing.assert_allclose(hg, ho.numpy(), atol=ATOL, rtol=0)
            np.testing.assert_allclose(lg, lo.numpy(), atol=ATOL, rtol=0)


def test_prefix_kv_reuse_matches_full_prefill(models):
    """clear_slow_caches_until keeps the conditioning KV (speech.rs:40): prefill of the
    remainder on top of a kept prefix == one full prefill."""
    cfg, tok, w, gpu, ora = models
    prompt = synth.make_prompt(cfg, tok, 40, seed=9)
    gpu.clear_slow_layer_caches()
    l_full, h_full = gpu.forward_generate(prompt[None], 0)
    gpu.clear_slow_caches_until(25)
    assert gpu.curr_kv_size() == 25
    l_part, h_part = gpu.forward_generate(prompt[None, :, 25:], 25)
    np.testing.assert_allclose(h_part, h_full, atol=1e-5, rtol=0)
    ora.clear_slow_layer_caches()
    with torch.no_grad():
        ora.forward_generate(t64(prompt[:, :25])[None], 0)
        lo, ho = ora.forward_generate(t64(prompt[:, 25:])[None], 25)
    np.testing.assert_allclose(h_part, ho.numpy(), atol=ATOL, rtol=0)
    np.testing.assert_allclose(l_part, lo.numpy(), atol=ATOL, rtol=0)


def test_fast_stack_matches_oracle(models):
    cfg, tok, w, gpu, ora = models
    rng = np.random.default_rng(3)
    ora.clear_fast_layer_caches()
    gpu.clear_fast_layer_caches()
    x = rng.standard_normal((1, 1, cfg["dim"])).astype(np.float32)
    for cb in range(cfg["num_codebooks"]):
        with torch.no_grad():
            lo = ora.forward_generate_fast(torch.from_numpy(x), cb)
        lg = gpu.forward_generate_fast(x, cb)
        np.testing.assert_allclose(lg, lo.numpy(), atol=ATOL, rtol=0)
        a = int(lo.argmax())
        e = gpu.fast_embeddings([a])
        np.testing.assert_array_equal(e[0], w["fast_embeddings.weight"][a].numpy())
        x = e.reshape(1, 1, -1)


@pytest.mark.parametrize("P,n_frames", [(24, 12), (166, 6)])
def test_greedy_generate_tokens_bit_exact(models, P, n_frames):
    cfg, tok, w, gpu, ora = models
    prompt = synth.make_prompt(cfg, tok, P, seed=1000)
    args = SamplingArgs(temp=0.0, repetition_penalty=1.4)
    got = generate_blocking(gpu, prompt, 400, args, fixed_len=n_frames)
    ora.clear_slow_layer_caches()
    with torch.no_grad():
        exp = ogen.generate_blocking(ora, t64(prompt), 400, osamp.SamplingArgs(temp=0.0, repetition_penalty=1.4),
                                     fixed_len=n_frames)
    np.testing.assert_array_equal(got.astype(np.int64), exp.numpy())
    assert gpu.curr_kv_size() == P + n_frames - 1


def test_greedy_generate_with_eos_and_budget(models):
    """EOS handling (Q4, single_batch.rs:153-156,262-266) and the max_new_tokens budget (Q3)."""
    cfg, tok, w, gpu, ora = models
    prompt = synth.make_prompt(cfg, tok, 20, seed=77)
    args = SamplingArgs(temp=0.0)
    for max_new in (20, 23, 30):
        got = generate_blocking(gpu, prompt, max_new, args)
        ora.clear_slow_layer_caches()
        with torch.no_grad():
            exp = ogen.generate_blocking(ora, t64(prompt), max_new, osamp.SamplingArgs(temp=0.0))
        np.testing.assert_array_equal(got.astype(np.int64), exp.numpy())


def test_sampled_generate_matches_oracle(models):
    """temp 0.7 / top_p 0.8 / top_k 256 / rep-pen 1.4 with the shared Philox stream: identical ids
    unless a draw lands within float rounding of a CDF boundary (none with this seed)."""
    cfg, tok, w, gpu, ora = models
    prompt = synth.make_prompt(cfg, tok, 32, seed=1001)
    got = generate_blocking(gpu, prompt, 400, SamplingArgs(0.7, 0.8, 256, 1.4, seed=11), fixed_len=10)
    ora.clear_slow_layer_caches()
    with torch.no_grad():
        exp = ogen.generate_blocking(ora, t64(prompt), 400, osamp.SamplingArgs(0.7, 0.8, 256, 1.4, seed=11),
                                     fixed_len=10)
    np.testing.assert_array_equal(got.astype(np.int64), exp.numpy())
    assert len(np.unique(got)) > 20  # really sampling, not collapsing


def test_static_batch_is_independent_rows(models):
    """Row i of the batched call == the bs=1 path on prompt i (ragged prompts, per-row KV)."""
    cfg, tok, w, gpu, ora = models
    prompts = [synth.make_prompt(cfg, tok, P, seed=1000 + i) for i, P in enumerate((24, 61, 40))]
    args = SamplingArgs(temp=0.0)
    got = generate_static_batch(gpu, prompts, 400, args, fixed_len=8)
    with torch.no_grad():
        exp = ogen.generate_independent_batch(ora, [t64(p) for p in prompts], 400, osamp.SamplingArgs(temp=0.0),
                                              fixed_len=8)
    for g, e in zip(got, exp):
        np.testing.assert_array_equal(g.astype(np.int64), e.numpy())


@pytest.mark.parametrize("mode", [1, 2])
def test_decode_modes_agree_with_oracle(tiny_lm, mode):
    """decode_mode 1 (per-op kernels + CUDA graph) and 2 (persistent megakernel) both match the oracle,
    greedy and sampled, single row and ragged batch."""
    cfg, tok, w = tiny_lm
    gpu = DualARTransformer(w, cfg, tok, max_batch=3, max_seq_len=256, decode_mode=mode)
    ora = oracle_model(cfg, tok, w)
    prompts = [synth.make_prompt(cfg, tok, P, seed=50 + i) for i, P in enumerate((30, 17, 44))]
    for sa, so in ((SamplingArgs(temp=0.0), osamp.SamplingArgs(temp=0.0)),
                   (SamplingArgs(0.7, 0.8, 256, 1.4, seed=5), osamp.SamplingArgs(0.7, 0.8, 256, 1.4, seed=5))):
        got = generate_static_batch(gpu, prompts, 400, sa, fixed_len=9)
        with torch.no_grad():
            exp = ogen.generate_independent_batch(ora, [t64(p) for p in prompts], 400, so, fixed_len=9)
        for g, e in zip(got, exp):
            np.testing.assert_array_equal(g.astype(np.int64), e.numpy())
    gpu.close()


@pytest.mark.parametrize("mode,dtype", [(0, "f32"), (1, "f32"), (1, "bf16")])
def test_batch_above_eight_rows(tiny_lm, mode, dtype):
    """11 ragged rows.  f32: the per-op path runs two GEMV groups (auto mode falls back to it above one
    megakernel group).  bf16: the block projections run on tcgen05 with the batch as the MMA N dimension.
    Every row must equal its bs=1 oracle generation (Philox row index = global row)."""
    cfg, tok, w = tiny_lm
    if dtype == "bf16":
        w = synth.make_lm_weights(cfg, seed=1234, round_bf16=True)
    gpu = DualARTransformer(w, cfg, tok, max_batch=11, max_seq_len=128, decode_mode=mode, dtype=dtype)
    ora = oracle_model(cfg, tok, w)
    prompts = [synth.make_prompt(cfg, tok, 12 + 5 * i, seed=300 + i) for i in range(11)]
    sa, so = SamplingArgs(0.7, 0.8, 256, 1.4, seed=9), osamp.SamplingArgs(0.7, 0.8, 256, 1.4, seed=9)
    got = generate_static_batch(gpu, prompts, 400, sa, fixed_len=5)
    with torch.no_grad():
        exp = ogen.generate_independent_batch(ora, [t64(p) for p in prompts], 400, so, fixed_len=5)
    for g, e in zip(got, exp):
        np.testing.assert_array_equal(g.astype(np.int64), e.numpy())
    gpu.close()


def test_bf16_weights_mode(tiny_lm):
    """weight_dtype bf16: weights stored bf16, fp32 math == oracle run on bf16-rounded weights."""
    cfg, tok, _ = tiny_lm
    w = synth.make_lm_weights(cfg, seed=1234, round_bf16=True)
    gpu = DualARTransformer(w, cfg, tok, dtype="bf16", max_seq_len=256)
    ora = oracle_model(cfg, tok, w)
    prompt = synth.make_prompt(cfg, tok, 30, seed=4)
    with torch.no_grad():
        lo, ho = ora.forward_generate(t64(prompt)[None], 0)
    lg, hg = gpu.forward_generate(prompt[None], 0)
    np.testing.assert_allclose(hg, ho.numpy(), atol=ATOL, rtol=0)
    np.testing.assert_allclose(lg, lo.numpy(), atol=ATOL, rtol=0)
    got = generate_blocking(gpu, prompt, 400, SamplingArgs(temp=0.0), fixed_len=8)
    ora.clear_slow_layer_caches()
    with torch.no_grad():
        exp = ogen.generate_blocking(ora, t64(prompt), 400, osamp.SamplingArgs(temp=0.0), fixed_len=8)
    np.testing.assert_array_equal(got.astype(np.int64), exp.numpy())
    gpu.close()


@pytest.mark.parametrize("P", [30, 200, 300])  # N tiles of 32, 64 and 128 prompt positions
def test_tcgen05_prefill_matches_fma_prefill_and_oracle(tiny_lm, P, monkeypatch):
    """bf16-weight prefill runs the dense projections on tcgen05 (TMA + TMEM, activations split into three
    bf16 terms): same hidden state / logits as the CUDA-core FMA prefill (1e-4) and as the oracle (1e-3)."""
    cfg, tok, _ = tiny_lm
    w = synth.make_lm_weights(cfg, seed=1234, round_bf16=True)
    prompt = synth.make_prompt(cfg, tok, P, seed=8)
    tc = DualARTransformer(w, cfg, tok, dtype="bf16", max_seq_len=512)
    l_tc, h_tc = tc.forward_generate(prompt[None], 0)
    tc.close()
    monkeypatch.setenv("FSB_NO_TCGEN05", "1")
    fma = DualARTransformer(w, cfg, tok, dtype="bf16", max_seq_len=512)
    l_fma, h_fma = fma.forward_generate(prompt[None], 0)
    fma.close()
    assert not np.array_equal(h_tc, h_fma)  # really two different code paths
    np.testing.assert_allclose(h_tc, h_fma, atol=1e-4, rtol=0)
    np.testing.assert_allclose(l_tc, l_fma, atol=1e-4, rtol=0)
    ora = oracle_model(cfg, tok, w)
    with torch.no_grad():
        lo, ho = ora.forward_generate(t64(prompt)[None], 0)
    np.testing.assert_allclose(h_tc, ho.numpy(), atol=ATOL, rtol=0)
    np.testing.assert_allclose(l_tc, lo.numpy(), atol=ATOL, rtol=0)


def test_fish14_legacy_slow_token(tiny_lm):
    """Fish <= 1.4 (Q8): slow token is PAD until the budget ends (fixed_len), fast codes greedy."""
    cfg, _, w = tiny_lm
    tok = dict(im_end_id=4, pad_id=5, semantic_start_id=5, semantic_end_id=None)
    gpu = DualARTransformer(w, cfg, tok, fish_version="1.4", max_seq_len=256)
    ora = oracle_model(cfg, tok, w, "1.4")
    prompt = synth.make_prompt(cfg, tok, 28, seed=2)
    got = generate_blocking(gpu, prompt, 400, SamplingArgs(temp=0.0), fixed_len=6)
    with torch.no_grad():
        exp = ogen.generate_blocking(ora, t64(prompt), 400, osamp.SamplingArgs(temp=0.0), fixed_len=6,
                                     force_slow=[tok["pad_id"]])
    np.testing.assert_array_equal(got.astype(np.int64), exp.numpy())
    gpu.close()


def test_errors_are_reported(models):
    cfg, tok, w, gpu, ora = models
    from fish_speech_rs_b200._ffi import FsbError
    with pytest.raises(FsbError):  # KV overflow
        generate_blocking(gpu, synth.make_prompt(cfg, tok, 500, seed=1), 2000, SamplingArgs(temp=0.0), fixed_len=100)
    bad = dict(w)
    del bad["norm.weight"]
    with pytest.raises(FsbError) as e:
        DualARTransformer(bad, cfg, tok)
    assert e.value.status == -3 and "norm.weight" in str(e.value)


@pytest.mark.parametrize("dtype,sync", [("f32", "barrier"), ("bf16", "barrier"), ("bf16", "ll")])
def test_single_row_ring_megakernel_full_width(dtype, sync, monkeypatch):
    """Full-width blocks (dim 1024 / FFN 4096 / 16 q + 2 kv heads): one row takes the single-row megakernel
    (TMA weight ring, register-resident activations, folded RMSNorm).  Token ids equal the oracle's, greedy
    and sampled, over enough frames to cross an attention-chunk boundary (64 positions) and recycle the ring."""
    if sync == "ll":  # flag-in-data synchronisation instead of grid barriers (opt-in variant of the same kernel)
        monkeypatch.setenv("FSB_MEGA_LL", "1")
    cfg, tok = dict(synth.WIDE), dict(synth.TINY_TOKENS)
    w = synth.make_lm_weights(cfg, seed=77, round_bf16=(dtype == "bf16"))
    gpu = DualARTransformer(w, cfg, tok, max_batch=1, max_seq_len=256, decode_mode=2, dtype=dtype)
    ora = oracle_model(cfg, tok, w)
    prompt = synth.make_prompt(cfg, tok, 57, seed=21)
    for sa, so in ((SamplingArgs(temp=0.0), osamp.SamplingArgs(temp=0.0)),
                   (SamplingArgs(0.7, 0.8, 256, 1.4, seed=11), osamp.SamplingArgs(0.7, 0.8, 256, 1.4, seed=11))):
        got = generate_blocking(gpu, prompt, 400, sa, fixed_len=12)
        ora.clear_slow_layer_caches()
        with torch.no_grad():
            exp = ogen.generate_blocking(ora, t64(prompt), 400, so, fixed_len=12)
        np.testing.assert_array_equal(got.astype(np.int64), exp.numpy())
    # natural stop (no fixed length): the EOS / budget path of the frame loop, and a second call on the same handle
    got = generate_blocking(gpu, prompt, 57 + 6, SamplingArgs(temp=0.0))
    ora.clear_slow_layer_caches()
    with torch.no_grad():
        exp = ogen.generate_blocking(ora, t64(prompt), 57 + 6, osamp.SamplingArgs(temp=0.0))
    np.testing.assert_array_equal(got.astype(np.int64), exp.numpy())
    gpu.close()


def test_cfg4_voice_clone_chain(tiny_lm, codec_weights):
    """cfg4 plumbing: wav -> log-mel -> codes on the GPU (`FireflyCodec::encode`), codes (+1, prompt.rs:88) into a
    Fish-1.4 style prompt, conditioned greedy decode, vocode.  Same prompt through the oracle: same tokens, same PCM."""
    from fish_speech_rs_b200 import FireflyCodec
    from oracle import codec as ocodec
    cfg, _, w = tiny_lm
    tok = dict(im_end_id=4, pad_id=5, semantic_start_id=5, semantic_end_id=None)
    codec = FireflyCodec(codec_weights, max_frames=64, with_encoder=True)
    rng = np.random.default_rng(11)
    pcm = (0.2 * rng.standard_normal(2048 * 12)).astype(np.float32)
    ref_codes = codec.encode(pcm)[0]  # (8, L)
    L = ref_codes.shape[1]
    assert L == ((((2048 * 12 + 1536) // 512 - 3) - 2) // 2 + 1 - 2) // 2 + 1
    prompt = synth.make_prompt(cfg, tok, L + 12, seed=3, voice=(ref_codes + 1).astype(np.int64))
    gpu = DualARTransformer(w, cfg, tok, fish_version="1.4", max_seq_len=256)
    ora = oracle_model(cfg, tok, w, "1.4")
    got = generate_blocking(gpu, prompt, 400, SamplingArgs(temp=0.0), fixed_len=5)
    with torch.no_grad():
        exp = ogen.generate_blocking(ora, t64(prompt), 400, osamp.SamplingArgs(temp=0.0), fixed_len=5,
                                     force_slow=[tok["pad_id"]])
    np.testing.assert_array_equal(got.astype(np.int64), exp.numpy())
    out_codes = np.minimum(got, 999)[None]  # random-weight LMs emit codes >= 1000 (Q11): harness-side clamp
    wav = codec.decode(out_codes)
    with torch.no_grad():
        wav_o = ocodec.decode(t64(out_codes), codec_weights).numpy()
    np.testing.assert_allclose(wav, wav_o, atol=1e-4, rtol=0)
    gpu.close()
    codec.close()


def test_full_size_properties_fish15():
    """BASELINE.json's full Fish-1.5 shapes (24 + 4 layers, vocab 102 048; too big for the CPU oracle): properties that
    do not need it.  (i) the single-row megakernel and the per-op CUDA-graph path -- two independent implementations of
    the frame loop, with different samplers (radix select vs bitonic sort) -- emit the same codes, greedy and sampled;
    (ii) a second call with the same seed repeats them, another seed does not; (iii) exactly fixed_len frames, codes < 1024."""
    cfg, tok = dict(synth.FISH15), dict(synth.FISH15_TOKENS)
    w = synth.make_lm_weights(cfg, seed=1234, round_bf16=True)
    voice = np.load(os.path.join(os.path.dirname(__file__), "golden", "default_voice.npy"))
    prompt = synth.make_prompt(cfg, tok, 384, seed=1000, voice=voice)
    outs = {}
    for mode in (2, 1):
        lm = DualARTransformer(w, cfg, tok, dtype="bf16", max_batch=1, max_seq_len=448, decode_mode=mode)
        outs[mode, "greedy"] = generate_blocking(lm, prompt, 100000, SamplingArgs(temp=0.0), fixed_len=10)
        outs[mode, "s5"] = generate_blocking(lm, prompt, 100000, SamplingArgs(0.7, 0.8, 256, 1.4, seed=5), fixed_len=10)
        if mode == 2:
            outs["again"] = generate_blocking(lm, prompt, 100000, SamplingArgs(0.7, 0.8, 256, 1.4, seed=5), fixed_len=10)
            outs["s6"] = generate_blocking(lm, prompt, 100000, SamplingArgs(0.7, 0.8, 256, 1.4, seed=6), fixed_len=10)
        lm.close()
    for k in ("greedy", "s5"):
        assert outs[2, k].shape == (8, 10) and outs[2, k].max() < 1024
        np.testing.assert_array_equal(outs[2, k], outs[1, k])
    np.testing.assert_array_equal(outs["again"], outs[2, "s5"])
    assert not np.array_equal(outs["s6"], outs[2, "s5"])


# ----------------------------------------------------------------------------------------------------------------
# Wide batches (cfg3 / cfg5 path): one persistent tcgen05 + TMA megakernel for 9..32 rows (fsb_lm_megab.cuh).
# A tensor-core path sums in a different order than the CPU reference, so a handful of near-ties flip over thousands of
# decisions on flat synthetic distributions.  The check is therefore the teacher-forced replay of oracle/generate.py:
# EVERY sampling decision of EVERY row is re-derived by the oracle from the same history and must either be identical
# or a proven near-tie (oracle margin below the fp tolerance: logits 2e-3, CDF 2e-3 / temp); zero violations allowed.
def replay_all(gpu, ora, prompts, outs, so, fixed_len, force_slow=False):
    tot = dict(decisions=0, exact=0, near_tie=0, violation=0)
    events = []
    for i, p in enumerate(prompts):
        fr = gpu.last_frames(i)
        if fixed_len is not None:
            assert fr.shape[1] == fixed_len
            np.testing.assert_array_equal(fr[1:], outs[i])  # generate_* output == frames minus the semantic row
        r = ogen.replay_frames(ora, t64(p), fr, so, row=i, fixed_len=fixed_len, force_slow=force_slow)
        for k in tot:
            tot[k] += r[k]
        events += [(i,) + ev for ev in r["events"] if ev[-1] != "near_tie"]
    return tot, events


def wide_models(nrows, max_seq_len=256, decode_mode=2, seed=77):
    cfg, tok = dict(synth.WIDE), dict(synth.TINY_TOKENS)
    w = synth.make_lm_weights(cfg, seed=seed, round_bf16=True)
    gpu = DualARTransformer(w, cfg, tok, max_batch=nrows, max_seq_len=max_seq_len, decode_mode=decode_mode, dtype="bf16")
    return cfg, tok, w, gpu, oracle_model(cfg, tok, w)


@pytest.mark.parametrize("nrows,frames", [(11, 8), (16, 40), (27, 12)])  # NPAD 16, 16 (rep-pen window 16 evicts), 32
def test_wide_batch_megakernel_matches_oracle(nrows, frames):
    cfg, tok, w, gpu, ora = wide_models(nrows)
    prompts = [synth.make_prompt(cfg, tok, 12 + 7 * i, seed=300 + i) for i in range(nrows)]  # ragged: 12 .. 194
    for sa, so in ((SamplingArgs(temp=0.0), osamp.SamplingArgs(temp=0.0)),
                   (SamplingArgs(0.7, 0.8, 256, 1.4, seed=9), osamp.SamplingArgs(0.7, 0.8, 256, 1.4, seed=9))):
        outs = generate_static_batch(gpu, prompts, 400, sa, fixed_len=frames)
        assert gpu.stats()["kernel_launches"] < 60 * nrows  # prefill launches + ONE decode launch (not ~1450 per frame)
        tot, bad = replay_all(gpu, ora, prompts, outs, so, frames)
        assert tot["violation"] == 0, bad[:5]
        assert tot["decisions"] == nrows * frames * 9
        assert tot["exact"] >= 0.99 * tot["decisions"], tot
    gpu.close()


def test_wide_batch_equals_single_row_paths():
    """Row i of the 12-row launch must reproduce the single-row megakernel on the same prompt and Philox row -- two
    kernels with different summation orders, so compare through the replay as well as directly."""
    cfg, tok, w, gpu, ora = wide_models(12)
    prompts = [synth.make_prompt(cfg, tok, 20 + 9 * i, seed=40 + i) for i in range(12)]
    sa = SamplingArgs(temp=0.0)
    outs = generate_static_batch(gpu, prompts, 400, sa, fixed_len=10)
    one = DualARTransformer(w, cfg, tok, max_batch=1, max_seq_len=256, decode_mode=2, dtype="bf16")
    same = 0
    for i in (0, 5, 11):
        same += int(np.array_equal(generate_blocking(one, prompts[i], 400, sa, fixed_len=10), outs[i]))
    assert same >= 2  # greedy rows agree unless a near-tie flips (the replay above bounds those)
    one.close()
    gpu.close()


def test_ragged_batch_natural_stop_at_arena_end():
    """rows that exhaust their budget sit at pos == max_new_tokens + 1 == max_len while shorter prompts go on.
    They must neither append K/V (out of bounds) nor disturb live rows.  No fixed_len: rows stop at different frames."""
    max_len = 96
    for mode, nrows in ((1, 4), (2, 4), (2, 12)):  # per-op, 2-8 row megakernel, wide-batch megakernel
        cfg, tok, w, gpu, ora = wide_models(nrows, max_seq_len=max_len, decode_mode=mode)
        lens = [90 - 6 * i for i in range(nrows)]  # the longest prompt finishes first
        prompts = [synth.make_prompt(cfg, tok, P, seed=700 + i) for i, P in enumerate(lens)]
        sa, so = SamplingArgs(temp=0.0), osamp.SamplingArgs(temp=0.0)
        outs = generate_static_batch(gpu, prompts, max_len - 1, sa)
        tot, bad = replay_all(gpu, ora, prompts, outs, so, None)
        assert tot["violation"] == 0, (mode, nrows, bad[:5])
        for i, P in enumerate(lens):
            fr = gpu.last_frames(i)
            # Q3: frames stop once input_pos exceeds max_new_tokens (+ an earlier <|im_end|>)
            assert fr.shape[1] <= max_len - 1 - P + 2
            assert fr.shape[1] == max_len - 1 - P + 2 or fr[0, -1] == tok["im_end_id"]
        gpu.close()


def test_bad_arguments_are_rejected(models):
    """include/fsb.h promises FSB_ERR_INVALID (the reference errors too: WeightedIndex on an empty set, index_select
    out of range) instead of out-of-bounds device accesses."""
    cfg, tok, w, gpu, ora = models
    from fish_speech_rs_b200._ffi import FsbError
    prompt = synth.make_prompt(cfg, tok, 16, seed=1)
    for bad_args in (SamplingArgs(0.7, 0.8, 0, 1.4), SamplingArgs(float("nan"), 0.8, 256, 1.4),
                     SamplingArgs(0.7, 0.8, 256, 0.0)):
        with pytest.raises(FsbError) as e:
            generate_blocking(gpu, prompt, 64, bad_args, fixed_len=2)
        assert e.value.status == -1
    p2 = prompt.copy()
    p2[0, 3] = cfg["vocab_size"]
    with pytest.raises(FsbError) as e:
        generate_blocking(gpu, p2, 64, SamplingArgs(temp=0.0), fixed_len=2)
    assert e.value.status == -1 and "vocab_size" in str(e.value)
    p3 = prompt.copy()
    p3[4, 5] = cfg["codebook_size"]
    with pytest.raises(FsbError) as e:
        generate_blocking(gpu, p3, 64, SamplingArgs(temp=0.0), fixed_len=2)
    assert e.value.status == -1 and "codebook_size" in str(e.value)
    # the handle is still usable
    assert generate_blocking(gpu, prompt, 64, SamplingArgs(temp=0.0), fixed_len=2).shape == (cfg["num_codebooks"], 2)


# ----------------------------------------------------------------------------------------------------------------
# Branches the round-1 tests never reached (VERDICT r1 "windows that never close")
NONADJ_TOKENS = dict(im_end_id=1100, pad_id=1090, semantic_start_id=1264, semantic_end_id=1264 + 1023)


@pytest.mark.parametrize("mode,nrows,dtype", [(1, 1, "f32"), (2, 1, "f32"), (2, 1, "bf16"), (2, 3, "f32"), (2, 10, "bf16")])
def test_non_adjacent_im_end_constrained_head(mode, nrows, dtype):
    """generate/utils.rs:17-33: when <|im_end|> is NOT directly in front of the semantic range the constrained logits
    are cat(logits[im_end], logits[semantic_start..]) and ids are rescaled back through the two-piece map.  Exercised on
    the per-op path, the single-row ring kernel, the 2-8 row kernel and the wide-batch kernel (extra head tile)."""
    cfg, tok = dict(synth.WIDE), dict(NONADJ_TOKENS)
    w = synth.make_lm_weights(cfg, seed=91, round_bf16=(dtype == "bf16"))
    gpu = DualARTransformer(w, cfg, tok, max_batch=nrows, max_seq_len=160, decode_mode=mode, dtype=dtype)
    ora = oracle_model(cfg, tok, w)
    prompts = [synth.make_prompt(cfg, tok, 20 + 6 * i, seed=900 + i) for i in range(nrows)]
    for sa, so in ((SamplingArgs(temp=0.0), osamp.SamplingArgs(temp=0.0)),
                   (SamplingArgs(0.7, 0.8, 256, 1.4, seed=2), osamp.SamplingArgs(0.7, 0.8, 256, 1.4, seed=2))):
        # natural stop: <|im_end|> (constrained index 0) is eligible and must be handled through the rescale map
        outs = generate_static_batch(gpu, prompts, 60, sa)
        tot, bad = replay_all(gpu, ora, prompts, outs, so, None)
        assert tot["violation"] == 0, bad[:5]
        assert tot["exact"] >= 0.98 * tot["decisions"]
        for i in range(nrows):
            fr = gpu.last_frames(i)
            sem = fr[0][fr[0] != tok["im_end_id"]]
            assert ((sem >= tok["semantic_start_id"]) & (sem <= tok["semantic_end_id"] + 16)).all()
    gpu.close()


def test_tied_word_embeddings(tiny_lm):
    """dual_ar.rs:486-490: `output` is the embedding table when tie_word_embeddings is set (no output.weight in the file)."""
    cfg, tok, w = tiny_lm
    cfg = dict(cfg, tie_word_embeddings=True)
    w = {k: v for k, v in w.items() if k != "output.weight"}
    gpu = DualARTransformer(w, cfg, tok, max_seq_len=128)
    ora = oracle_model(cfg, tok, w)
    prompt = synth.make_prompt(cfg, tok, 30, seed=4)
    with torch.no_grad():
        lo, ho = ora.forward_generate(t64(prompt)[None], 0)
    lg, hg = gpu.forward_generate(prompt[None], 0)
    np.testing.assert_allclose(lg, lo.numpy(), atol=ATOL, rtol=0)
    got = generate_blocking(gpu, prompt, 400, SamplingArgs(temp=0.0), fixed_len=8)
    ora.clear_slow_layer_caches()
    with torch.no_grad():
        exp = ogen.generate_blocking(ora, t64(prompt), 400, osamp.SamplingArgs(temp=0.0), fixed_len=8)
    np.testing.assert_array_equal(got.astype(np.int64), exp.numpy())
    gpu.close()


@pytest.mark.parametrize("mode,dtype", [(1, "f32"), (2, "f32"), (2, "bf16")])
def test_prefix_kv_reuse_two_chunk_generation(mode, dtype):
    """server/lib/handlers/speech.rs:27-40: chunk 0 = [conditioning | text 0], generate, `clear_slow_caches_until(n_cond)`,
    chunk 1 = [text 1] on top of the kept conditioning KV (FSB_GEN_KEEP_SLOW_KV).  Same calls on the oracle."""
    cfg, tok = dict(synth.WIDE), dict(synth.TINY_TOKENS)
    w = synth.make_lm_weights(cfg, seed=33, round_bf16=(dtype == "bf16"))
    gpu = DualARTransformer(w, cfg, tok, max_batch=1, max_seq_len=256, decode_mode=mode, dtype=dtype)
    ora = oracle_model(cfg, tok, w)
    n_cond = 70
    full = synth.make_prompt(cfg, tok, n_cond + 14, seed=61)
    chunk1 = synth.make_prompt(cfg, tok, 9, seed=62)  # text only (P < voice span): continues after the conditioning
    sa, so = SamplingArgs(0.7, 0.8, 256, 1.4, seed=8), osamp.SamplingArgs(0.7, 0.8, 256, 1.4, seed=8)
    got0 = generate_blocking(gpu, full, 400, sa, fixed_len=7)
    assert gpu.curr_kv_size() == n_cond + 14 + 6
    fr0 = gpu.last_frames(0)
    gpu.clear_slow_caches_until(n_cond)
    assert gpu.curr_kv_size() == n_cond
    got1 = generate_blocking(gpu, chunk1, 400, sa, fixed_len=7, keep_slow_kv=True)
    fr1 = gpu.last_frames(0)
    assert gpu.curr_kv_size() == n_cond + 9 + 6
    # oracle: replay chunk 0, truncate, replay chunk 1 on the kept prefix
    r0 = ogen.replay_frames(ora, t64(full), fr0, so, row=0, fixed_len=7)
    assert r0["violation"] == 0 and r0["exact"] >= r0["decisions"] - 2
    ora.clear_slow_layer_caches()
    with torch.no_grad():
        ora.forward_generate(t64(full[:, :n_cond])[None], 0)  # the conditioning KV the server keeps
    r1 = ogen.replay_frames(ora, t64(chunk1), fr1, so, row=0, fixed_len=7, keep_slow_kv=True)
    assert r1["violation"] == 0 and r1["exact"] >= r1["decisions"] - 2
    assert got0.shape == got1.shape == (8, 7) and not np.array_equal(got0, got1)
    gpu.close()


def test_full_size_fish15_against_the_oracle():
    """BASELINE shapes (24 + 4 layers, dim 1024, vocab 102 048), bf16-rounded weights, P = 384 with the default voice:
    24 frames of the single-row ring kernel and 6 frames of a 9-row wide-batch launch, every sampling decision replayed by
    the oracle (40 ms per frame on the CPU -- the round-1 claim that this was too big for the oracle was wrong)."""
    cfg, tok = dict(synth.FISH15), dict(synth.FISH15_TOKENS)
    w = synth.make_lm_weights(cfg, seed=1234, round_bf16=True)
    ora = oracle_model(cfg, tok, w)
    voice = np.load(os.path.join(os.path.dirname(__file__), "golden", "default_voice.npy"))
    prompt = synth.make_prompt(cfg, tok, 384, seed=1000, voice=voice)
    lm = DualARTransformer(w, cfg, tok, dtype="bf16", max_batch=9, max_seq_len=448, decode_mode=2)
    for sa, so in ((SamplingArgs(temp=0.0), osamp.SamplingArgs(temp=0.0)),
                   (SamplingArgs(0.7, 0.8, 256, 1.4, seed=5), osamp.SamplingArgs(0.7, 0.8, 256, 1.4, seed=5))):
        out = generate_blocking(lm, prompt, 100000, sa, fixed_len=24)
        r = ogen.replay_frames(ora, t64(prompt), lm.last_frames(0), so, row=0, fixed_len=24)
        assert r["violation"] == 0 and r["decisions"] == 24 * 9 and r["exact"] >= r["decisions"] - 3, r
        assert out.shape == (8, 24)
    prompts = [synth.make_prompt(cfg, tok, 300 + 11 * i, seed=2000 + i, voice=voice) for i in range(9)]
    sa, so = SamplingArgs(0.7, 0.8, 256, 1.4, seed=6), osamp.SamplingArgs(0.7, 0.8, 256, 1.4, seed=6)
    outs = generate_static_batch(lm, prompts, 100000, sa, fixed_len=6)
    tot, bad = replay_all(lm, ora, prompts, outs, so, 6)
    assert tot["violation"] == 0 and tot["exact"] >= tot["decisions"] - 5, (tot, bad[:5])
    lm.close()


def test_generate_blocking_with_hidden(models):
    """single_batch.rs:217-306 with collect_hidden_states: one pre-norm slow hidden state per yielded frame
    (<|im_end|> frames included), equal to what the oracle's iterator saw (atol 1e-3), codes unchanged."""
    from fish_speech_rs_b200 import generate_blocking_with_hidden
    cfg, tok, w, gpu, ora = models
    prompt = synth.make_prompt(cfg, tok, 26, seed=314)
    for max_new, fixed in ((400, 9), (30, None)):
        codes, hid = generate_blocking_with_hidden(gpu, prompt, max_new, SamplingArgs(temp=0.0), fixed_len=fixed)
        np.testing.assert_array_equal(codes, generate_blocking(gpu, prompt, max_new, SamplingArgs(temp=0.0), fixed_len=fixed))
        ora.clear_slow_layer_caches()
        gen = ogen.SingleBatchGenerator(ora, t64(prompt), max_new, osamp.SamplingArgs(temp=0.0), True, 0, fixed)
        exp = []
        with torch.no_grad():
            while True:
                if fixed is not None and len(exp) >= fixed:
                    break
                fr = gen.next()
                if fr is None:
                    break
                exp.append(gen.last_hidden.reshape(1, -1).numpy().copy())
        ora.clear_slow_layer_caches()
        assert hid.shape == (len(exp), 1, cfg["dim"])
        np.testing.assert_allclose(hid[:, 0], np.concatenate(exp, 0), atol=ATOL, rtol=0)


def test_kv_snapshot_restores_the_voice_prefix(tiny_lm):
    """the conditioning (system + voice) KV is prefilled once, snapshotted on the device and restored into
    another row before an utterance; generation of the remaining columns on top of it (FSB_GEN_KEEP_SLOW_KV, the server's
    `clear_slow_caches_until` flow, speech.rs:40) must equal the oracle's generation from the full prompt."""
    cfg, tok, w = tiny_lm
    gpu = DualARTransformer(w, cfg, tok, max_batch=3, max_seq_len=256)
    ora = oracle_model(cfg, tok, w)
    n_cond = 48
    full = synth.make_prompt(cfg, tok, n_cond + 11, seed=808)
    gpu.forward_generate(full[None, :, :n_cond], 0, want_logits=False)  # prefill of the conditioning on row 0
    assert gpu.curr_kv_size() == n_cond
    snap = gpu.kv_snapshot_save(0, n_cond)
    gpu.clear_slow_layer_caches()
    # (another voice / utterance trashes row 0 in between)
    generate_blocking(gpu, synth.make_prompt(cfg, tok, 70, seed=1), 400, SamplingArgs(temp=0.0), fixed_len=3)
    gpu.kv_snapshot_restore(snap, 0)
    assert gpu.curr_kv_size() == n_cond
    sa, so = SamplingArgs(0.7, 0.8, 256, 1.4, seed=4), osamp.SamplingArgs(0.7, 0.8, 256, 1.4, seed=4)
    got = generate_blocking(gpu, full[:, n_cond:], 400, sa, fixed_len=9, keep_slow_kv=True)
    assert gpu.curr_kv_size() == n_cond + 11 + 8
    ora.clear_slow_layer_caches()
    with torch.no_grad():
        exp = ogen.generate_blocking(ora, t64(full), 400, so, fixed_len=9)
    np.testing.assert_array_equal(got.astype(np.int64), exp.numpy())
    gpu.kv_snapshot_free(snap)
    gpu.close()


@pytest.mark.parametrize("fixed", [True, False])
def test_continuous_batching_session(fixed):
    """13 utterances through 9 slots.  A slot is refilled as soon as its utterance ends while the other
    slots keep decoding (the reference's static batch waits for the slowest row, static_batch.rs:160-173; its server
    serialises whole generations, state.rs:13).  Every utterance is checked decision by decision against the oracle
    (teacher-forced replay, Philox row == slot) and must have the length the bs=1 rules give it."""
    nslots, nutt = 9, 13
    cfg, tok, w, gpu, ora = wide_models(nslots, max_seq_len=160)
    rng = np.random.default_rng(5)
    prompts = [synth.make_prompt(cfg, tok, int(rng.integers(14, 60)), seed=4000 + i) for i in range(nutt)]
    lens = [int(rng.integers(3, 19)) for _ in range(nutt)]  # fixed: exact frame counts; else: max_new_tokens budgets
    sa, so = SamplingArgs(0.7, 0.8, 256, 1.4, seed=21), osamp.SamplingArgs(0.7, 0.8, 256, 1.4, seed=21)
    gpu.session_begin(sa, fixed_len=fixed)
    slot_utt = {}
    results, frames_of, slot_of = {}, {}, {}
    nxt = 0

    def admit(slot):
        nonlocal nxt
        i = nxt
        nxt += 1
        P = prompts[i].shape[1]
        gpu.session_admit(slot, prompts[i], 400 if fixed else P + lens[i] - 2, fixed_len=lens[i] if fixed else 0)
        slot_utt[slot] = i
        slot_of[i] = slot

    for s in range(nslots):
        admit(s)
    launches = 0
    while slot_utt:
        act = gpu.session_run(4)
        launches += 1
        assert launches < 40
        for s in list(slot_utt):
            if not act[s]:
                i = slot_utt.pop(s)
                frames_of[i] = gpu.last_frames(s)
                results[i] = gpu.session_collect(s, 64)
                if nxt < nutt:
                    admit(s)
    assert sorted(results) == list(range(nutt))
    tot = dict(decisions=0, exact=0, near_tie=0, violation=0)
    for i in range(nutt):
        fr = frames_of[i]
        np.testing.assert_array_equal(results[i], fr[1:, [f for f in range(fr.shape[1]) if f == 0 or fr[0, f] != tok["im_end_id"]]])
        if fixed:
            assert fr.shape[1] == lens[i]
        else:
            assert fr.shape[1] == lens[i] or fr[0, -1] == tok["im_end_id"]  # Q3 budget or an earlier <|im_end|>
        r = ogen.replay_frames(ora, t64(prompts[i]), fr, so, row=slot_of[i], fixed_len=lens[i] if fixed else None)
        assert r["violation"] == 0, (i, r["events"][:4])
        for k in tot:
            tot[k] += r[k]
    assert tot["exact"] >= 0.98 * tot["decisions"], tot
    gpu.close()


def test_sharded_synthesizer_drives_the_library(tiny_lm):
    """fish_speech_rs_b200.shard.ShardedSynthesizer: two ranks emulated one after the other on this GPU; each
    runs only its launches through generate_static_batch, greedy rows equal the oracle's single-utterance output whatever
    launch and position they landed in, and the two shares cover the job exactly once."""
    from fish_speech_rs_b200 import shard
    cfg, tok, w = tiny_lm
    gpu = DualARTransformer(w, cfg, tok, max_batch=3, max_seq_len=256)
    ora = oracle_model(cfg, tok, w)
    prompts = [synth.make_prompt(cfg, tok, P, seed=700 + i) for i, P in enumerate((21, 48, 33, 64, 17, 40, 55))]
    got = {}
    for r in range(2):
        syn = shard.ShardedSynthesizer(lm=gpu, rank=r, world_size=2, sampling_args=SamplingArgs(temp=0.0), max_rows=3,
                                       vocode=lambda codes: [c.shape[1] for c in codes])
        res = syn.synthesize(prompts, 400, fixed_len=6)
        assert all(len(b) <= 3 for b in syn.launches) and not set(res) & set(got)
        got.update(res)
    assert sorted(got) == list(range(len(prompts)))
    with torch.no_grad():  # (clears the oracle's slow KV before every row)
        exp = ogen.generate_independent_batch(ora, [t64(p) for p in prompts], 400, osamp.SamplingArgs(temp=0.0), fixed_len=6)
    for i in range(len(prompts)):
        np.testing.assert_array_equal(got[i][0].astype(np.int64), exp[i].numpy())
        assert got[i][1] == 6
    gpu.close()

"""CPU tests pinning the oracle on the closed-form known answers the reference
carries.  The reference has no golden vectors for the LM / codec
arithmetic ("parity unpinned"), so these are the pins that exist."""
import os

import numpy as np
import pytest
import torch

from oracle import codec as ocodec
from oracle import dual_ar as olm
from oracle import generate as ogen
from oracle import rng as orng
from oracle import sampling as osamp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert orng.philox4x32_10((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    assert orng.philox4x32_10((f, f, f, f), (f, f)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert orng.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    u = orng.philox_uniform(7, 3, 1)
    assert 0.0 <= float(u) < 1.0


def test_fsq_implicit_codebook_closed_form():
    """fsq.rs:132-159: index -> digits base (8,5,5,5), basis (1,8,40,200), (d - half)/half."""
    cb = ocodec.fsq_implicit_codebook().numpy()
    assert cb.shape == (1000, 4)
    for idx in (0, 1, 7, 8, 39, 40, 199, 200, 999, 537):
        d = [idx % 8, (idx // 8) % 5, (idx // 40) % 5, (idx // 200) % 5]
        exp = [(d[0] - 4) / 4, (d[1] - 2) / 2, (d[2] - 2) / 2, (d[3] - 2) / 2]
        np.testing.assert_allclose(cb[idx], exp, rtol=0, atol=0)
    assert len({tuple(r) for r in cb.tolist()}) == 1000  # a bijection


def test_fsq_encode_inverts_decode():
    """codes_to_indices(indices_to_codes(i)) == i (fsq.rs:118-140) through the bound/round path
    when project_in is the identity on the 4 code dims."""
    cb = ocodec.fsq_implicit_codebook()
    levels = torch.tensor(ocodec.LEVELS, dtype=torch.float32)
    half_width = torch.floor(levels / 2.0)
    zhat = cb * half_width + half_width
    idx = (zhat * torch.tensor([1.0, 8.0, 40.0, 200.0])).sum(-1).to(torch.int64)
    assert torch.equal(idx, torch.arange(1000))


def test_mask_truth_table():
    """get_mask_abs (dual_ar.rs:702-712): 1 == MASK iff key j is in the future of query i."""
    m = olm.get_mask_abs(3, 5, 8192).numpy()  # 3 new queries on top of 2 cached rows
    exp = np.array([[0, 0, 0, 1, 1], [0, 0, 0, 0, 1], [0, 0, 0, 0, 0]], np.uint8)
    np.testing.assert_array_equal(m, exp)
    assert olm.get_mask_abs(1, 1, 8192).numpy().tolist() == [[0]]
    m = olm.get_mask_abs(4, 4, 8192).numpy()
    np.testing.assert_array_equal(m, np.triu(np.ones((4, 4), np.uint8), 1))


class _M:
    def __init__(self, tc, mt="1.5"):
        self.token_config, self.model_type = tc, mt


def test_constrain_rescale_round_trip():
    """generate/utils.rs:6-56."""
    tc = olm.TokenConfig(im_end_id=100011, pad_id=5, semantic_start_id=100012, semantic_end_id=101035)
    V = 102048
    logits = torch.arange(V, dtype=torch.float32)[None, None]
    c = ogen.constrain_probs_to_audio(logits, _M(tc))
    assert c.shape[-1] == V - tc.im_end_id and float(c[0, 0, 0]) == tc.im_end_id  # to the END of vocab (Q5)
    for shifted in (0, 1, 1024, V - tc.im_end_id - 1):
        assert ogen.rescale_semantic_token(shifted, _M(tc)) == int(c[0, 0, shifted])
    # non-adjacent tokenizer: [im_end | semantic_start ..)
    tc2 = olm.TokenConfig(im_end_id=4, pad_id=5, semantic_start_id=100, semantic_end_id=1123)
    c2 = ogen.constrain_probs_to_audio(logits[..., :2000], _M(tc2))
    assert c2.shape[-1] == 1 + 2000 - 100
    for shifted in (0, 1, 500):
        assert ogen.rescale_semantic_token(shifted, _M(tc2)) == int(c2[0, 0, shifted])
    # <= 1.4: untouched
    assert ogen.constrain_probs_to_audio(logits, _M(tc, "1.4")).shape[-1] == V


def test_rep_pen_window_bug_for_bug():
    """rep_pen.rs:37-65: divide regardless of sign; a token leaves the mask as soon as ANY
    occurrence of it drops out of the window."""
    p = osamp.RepPenProcessor(8, 3, 2.0)
    l = torch.tensor([4.0, -4.0, 2.0, 2.0, 2.0, 2.0, 2.0, 2.0])
    out = p.apply(l, 1)
    assert out.tolist() == [4.0, -2.0, 2.0, 2.0, 2.0, 2.0, 2.0, 2.0]  # negative logit gets LESS negative
    p.apply(l, 2)
    p.apply(l, 1)  # window [1,2,1]
    out = p.apply(l, 3)  # window [3,1,2] after dropping the oldest 1 -> token 1 un-penalised although still inside
    assert out[1] == -4.0 and out[2] == 1.0 and out[3] == 1.0


def test_sampler_argmax_and_topk_topp():
    a = osamp.SamplingArgs(temp=0.0)
    assert osamp.sample(torch.tensor([0.1, 3.0, 3.0, -1.0]), a, 0) == 1  # first maximal index
    probs = np.array([0.5, 0.3, 0.15, 0.05], np.float32)
    # top_k=2 keeps {0,1}; top_p=0.4 is reached by the first entry alone -> always 0
    for u in (0.0, 0.3, 0.999):
        assert osamp.sample_from_probs(probs, 2, 0.4, np.float32(u)) == 0
    # top_p >= sum over top-k -> plain multinomial over the top-k weights
    assert osamp.sample_from_probs(probs, 2, 0.95, np.float32(0.1)) == 0
    assert osamp.sample_from_probs(probs, 2, 0.95, np.float32(0.9)) == 1


def test_default_voice_fixture():
    """tests/golden/default_voice.npy == voices-template/default.npy: int64 (8, 274), values 3..999;
    274 == code frames the encoder arithmetic yields for sky.wav (562 265 samples)."""
    v = np.load(os.path.join(GOLDEN, "default_voice.npy"))
    assert v.dtype == np.int64 and v.shape == (8, 274) and v.min() == 3 and v.max() == 999
    n = 562265
    mel_frames = (n + 2 * 768) // 512 + 1 - 4 + 1  # 1102 hop chunks, first output on the 4th
    assert mel_frames == 1099
    l1 = (mel_frames - 2) // 2 + 1
    assert ((l1 - 2) // 2 + 1) == 274


def test_rope_table_matches_closed_form():
    cfg = olm.BaseModelArgs()
    cos, sin = olm.precompute_freqs_cis(cfg)
    assert cos.shape == (8192, 32)
    assert float(cos[0, 0]) == 1.0 and float(sin[0, 5]) == 0.0
    np.testing.assert_allclose(float(cos[1, 0]), np.cos(1.0), rtol=1e-7)
    th = np.float32(1.0) / np.float32(np.float64(1e6) ** np.float64(np.float32(2.0 / 64)))
    np.testing.assert_allclose(float(sin[7, 1]), np.sin(np.float64(np.float32(7) * th)), rtol=1e-6)


def test_vocoder_shapes_and_causality(codec_weights):
    """Appendix B: (1,8,T) -> (1,1,2048 T); strictly causal (changing a late code leaves earlier audio untouched)."""
    rng = np.random.default_rng(7)
    codes = torch.from_numpy(rng.integers(0, 1000, size=(1, 8, 6)))
    with torch.no_grad():
        pcm = ocodec.decode(codes, codec_weights)
        assert pcm.shape == (1, 1, 2048 * 6)
        codes2 = codes.clone()
        codes2[0, :, 4] = (codes2[0, :, 4] + 17) % 1000
        pcm2 = ocodec.decode(codes2, codec_weights)
    assert torch.equal(pcm[..., : 2048 * 4], pcm2[..., : 2048 * 4])
    assert not torch.equal(pcm[..., 2048 * 4:], pcm2[..., 2048 * 4:])
    with pytest.raises(IndexError):
        ocodec.decode(torch.full((1, 8, 2), 1000), codec_weights)


def test_tiny_lm_generate_is_deterministic(tiny_lm):
    from fish_speech_rs_b200 import synth
    cfg, tok, w = tiny_lm
    m = olm.DualARTransformer(w, olm.BaseModelArgs(**cfg), olm.TokenConfig(**tok))
    prompt = torch.from_numpy(synth.make_prompt(cfg, tok, 24, seed=1000).astype(np.int64))
    a = osamp.SamplingArgs(temp=0.0)
    with torch.no_grad():
        o1 = ogen.generate_blocking(m, prompt, 40, a, fixed_len=4)
        m.clear_slow_layer_caches()
        o2 = ogen.generate_blocking(m, prompt, 40, a, fixed_len=4)
    assert o1.shape == (8, 4) and torch.equal(o1, o2)


# ---------------------------------------------------------------- log-mel front-end
def test_mel_filterbank_is_pinned_to_the_reference_table():
    """oracle.mel.mel_filterbank() was compared with the reference's embedded melfilters160.bytes where the reference
    tree is mounted (tests/golden/make_mel_golden.py); here: the recorded difference, and that the formula has not moved."""
    import hashlib
    import json
    import os

    from oracle import mel
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "mel_golden.json")))
    assert g["melfilters160_max_abs_diff"] < 5e-7
    fb = mel.mel_filterbank()
    assert fb.shape == (1025, 160) and fb.dtype == np.float32
    assert hashlib.sha256(fb.tobytes()).hexdigest() == g["oracle_fb_sha256"]
    assert int((fb > 0).sum()) == g["oracle_fb_nonzeros"]
    np.testing.assert_allclose(fb.sum(0)[:8], g["oracle_fb_col_sums_first8"], rtol=1e-6)


def test_mel_frame_count_and_code_frames_of_sky_wav():
    """562 265 samples (tests/resources/sky.wav) -> 1099 mel frames -> 274 code frames == voices-template/default.npy."""
    import json
    import os

    from oracle import mel
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "mel_golden.json")))["sky_wav"]
    assert mel.n_mel_frames(g["samples"]) == g["mel_frames"] == 1099
    lm = g["mel_frames"]
    l1 = (lm - 2) // 2 + 1
    assert (l1 - 2) // 2 + 1 == g["code_frames"] == 274
    voice = np.load(os.path.join(os.path.dirname(__file__), "golden", "default_voice.npy"))
    assert voice.shape == (8, 274)
    # chunk arithmetic of the streaming STFT: exact multiples of the hop emit no partial frame
    assert mel.n_mel_frames(512 * 10 - 1536) == 10 - 3
    assert mel.n_mel_frames(512 * 10 - 1536 + 1) == 10 - 3 + 1


def test_mel_of_a_pure_tone_peaks_in_the_right_filter():
    from oracle import mel
    sr, f0 = 44100, 1000.0
    t = np.arange(sr // 2) / sr
    m = mel.log_mel((0.5 * np.sin(2 * np.pi * f0 * t)).astype(np.float32))
    assert m.shape == (160, mel.n_mel_frames(len(t)))
    fb = mel.mel_filterbank()
    want = int(np.argmax(fb[int(round(f0 / (sr / 2) * 1024))]))
    assert abs(int(np.argmax(m[:, m.shape[1] // 2])) - want) <= 1
    assert m.min() >= np.log(1e-5) - 1e-6 and m.max() <= np.log(100.0) + 1e-6
    # reflect padding repeats the edge sample (spectrogram.rs:14-27)
    np.testing.assert_array_equal(mel.reflect_pad(np.arange(5, dtype=np.float32), 2), [1, 0, 0, 1, 2, 3, 4, 4, 3])


# ---------------------------------------------------------------- output stage (resample, s16, WAV header)
def test_output_stage_known_answers():
    from oracle import audio_out as ao
    x = np.array([0.0, 1.0, 0.0, -1.0], np.float32)
    # 2x upsampling: indices 0, .5, 1, 1.5, 2, 2.5, 3, 3.5 (ceil clamped to the last sample)
    np.testing.assert_allclose(ao.resample(x, 1, 2), [0, .5, 1, .5, 0, -.5, -1, -1], atol=0)
    np.testing.assert_array_equal(ao.resample(x, 44100, 44100), x)
    assert len(ao.resample(np.zeros(2048 * 3, np.float32), 44100, 24000)) == int(np.ceil(2048 * 3 * 24000 / 44100))
    # (x.clamp(-1, 1) * 32767) as i16: truncation toward zero, saturation, NaN -> 0
    np.testing.assert_array_equal(ao.to_i16(np.array([0.5, -0.5, 2.0, -2.0, 1e-5, np.nan, 0.99999], np.float32)),
                                  [16383, -16383, 32767, -32767, 0, 0, 32766])
    # write_pcm_as_wav (wav.rs:27-58): 44-byte header, RIFF length = 36 + 2 n, byte rate = 2 sr
    b = ao.wav_bytes(np.array([1, -2, 3], np.int16), 44100)
    assert len(b) == 44 + 6 and b[:4] == b"RIFF" and b[8:16] == b"WAVEfmt "
    assert int.from_bytes(b[4:8], "little") == 36 + 6 and int.from_bytes(b[28:32], "little") == 88200
    assert int.from_bytes(b[40:44], "little") == 6 and b[44:] == np.array([1, -2, 3], "<i2").tobytes()


# ---------------------------------------------------------------- loader ground truth held by the reference
def _read_dims(path):
    import re
    out = {}
    for ln in open(path):
        m = re.match(r"Name: (\S+), Shape: torch\.Size\(\[([0-9, ]*)\]\)", ln.strip())
        if m:
            out[m.group(1)] = tuple(int(x) for x in m.group(2).split(",") if x.strip())
    return out


def test_lm_weight_names_and_shapes_match_the_reference_dump():
    """docs/llama-weight-dict.txt (Fish 1.2 checkpoint: vocab 32 000, 4 codebooks) lists exactly the tensors
    `DualARTransformer::load` binds (dual_ar.rs:460-529); the synthetic checkpoint generator -- and with it the C loader,
    which resolves the same names -- must reproduce every name and shape."""
    from fish_speech_rs_b200 import synth
    ref = _read_dims(os.path.join(GOLDEN, "llama_weight_dict_fish12.txt"))
    assert len(ref) == 203
    cfg = dict(synth.FISH15, vocab_size=32000, num_codebooks=4, max_seq_len=4096)
    assert synth.lm_weight_shapes(cfg) == ref
    # Fish 1.5 only changes the two vocabulary-sized tables and the codebook count
    got15 = synth.lm_weight_shapes(dict(synth.FISH15))
    diff = {k for k in got15 if got15[k] != ref.get(k)}
    assert diff == {"embeddings.weight", "codebook_embeddings.weight", "output.weight"}


def test_codec_weight_names_and_shapes_match_the_reference_dump():
    """docs/weight-dims-default.txt is the Fish 1.2 generator: weight norm un-merged (`parametrizations.weight.original0/1`),
    no `.conv.` infix, 4 FSQ groups of 128.  Fish >= 1.4 (what csrc/ loads, codec/utils/mod.rs:25-40,83-95): merged weights
    under `<prefix>.conv.weight|bias`, 8 groups of 64 (codec/config.rs:155-168).  After that documented renaming every
    tensor of the dump must exist with the same shape, and nothing else may be generated."""
    import re
    from fish_speech_rs_b200 import synth
    ref = _read_dims(os.path.join(GOLDEN, "codec_weight_dims_fish12.txt"))
    assert len(ref) == 509
    wrapped = re.compile(r"^(head\.(conv_pre|conv_post|ups\.\d+|resblocks\.\d+\.blocks\.\d+\.convs[12]\.\d+)|"
                         r"quantizer\.(down|up)sample\.\d+\.0|.*\.dwconv|backbone\.downsample_layers\.0\.0)$")
    exp = {}
    for name, shape in ref.items():
        if name.endswith(".parametrizations.weight.original0"):
            assert shape[1:] == (1, 1)  # the weight-norm gain g, folded into the merged weight
            continue
        if name.endswith(".parametrizations.weight.original1"):
            base, leaf = name[: -len(".parametrizations.weight.original1")], "weight"
        else:
            base, leaf = name.rsplit(".", 1)
        if "residual_fsq.rvqs." in name:
            continue  # group count / width differ by config, checked below
        exp[(base + ".conv." + leaf) if wrapped.match(base) else (base + "." + leaf)] = shape
    got = synth.codec_weight_shapes(with_encoder=True)
    got_no_fsq = {k: v for k, v in got.items() if "residual_fsq.rvqs." not in k}
    # codec/config.rs:146-163: Fish 1.2 has ONE x2 down/up-sampling stage (downsample_factor [2]), >= 1.4 has two ([2, 2]):
    # the second stage is the only thing the >= 1.4 generator adds, and it repeats the first stage's shapes
    second = {k for k in got_no_fsq if re.match(r"quantizer\.(down|up)sample\.1\.", k)}
    assert set(got_no_fsq) - set(exp) == second and len(second) == 2 * 11
    for k in second:
        assert got_no_fsq[k] == exp[k.replace("sample.1.", "sample.0.")]
    assert {k: v for k, v in got_no_fsq.items() if k not in second} == exp
    # FSQ projections: 4 x (4 <- 128) in the 1.2 dump, 8 x (4 <- 64) for >= 1.4; same tensor set per group
    leaves = {"project_in.weight", "project_in.bias", "project_out.weight", "project_out.bias"}
    assert {k.split("rvqs.")[1].split(".", 1)[1] for k in ref if "rvqs." in k} == leaves
    assert {k.split("rvqs.")[1].split(".", 1)[1] for k in got if "rvqs." in k} == leaves
    assert ref["quantizer.residual_fsq.rvqs.3.project_in.weight"] == (4, 128)
    assert got["quantizer.residual_fsq.rvqs.7.project_in.weight"] == (4, 64)
    assert 4 * 128 == 8 * 64 == 512


def _sky_pcm():
    import wave
    with wave.open(os.path.join(GOLDEN, "sky.wav")) as w:
        assert w.getframerate() == 44100 and w.getnchannels() == 1 and w.getsampwidth() == 2
        raw = w.readframes(w.getnframes())
    return np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0


def test_sky_wav_through_the_oracle_front_end(codec_weights):
    """The reference's own fixture through the restated front-end: 562 265 samples -> 1099 log-mel frames -> 274 code
    frames, exactly the length of voices-template/default.npy (the voice the server ships for this very clip)."""
    from oracle import mel as omel
    pcm = _sky_pcm()
    assert pcm.shape == (562265,)
    m = omel.log_mel(pcm)
    assert m.shape == (160, 1099) and np.isfinite(m).all()
    assert m.min() >= np.log(1e-5) - 1e-6 and m.max() <= np.log(100.0) + 1e-6  # the clamp of spectrogram.rs:154-156
    with torch.no_grad():
        codes = ocodec.encode_mel(torch.from_numpy(m[None]), codec_weights)
    voice = np.load(os.path.join(GOLDEN, "default_voice.npy"))
    assert tuple(codes.shape) == (1, 8, 274) == (1,) + voice.shape
    assert 0 <= int(codes.min()) and int(codes.max()) < 1000

"""GPU parity tests of the Firefly codec against the oracle, through the C ABI.
Tolerance: PCM max-abs 1e-4 on the tanh-bounded output, fp32 both sides."""
import os

import numpy as np
import pytest
import torch

from fish_speech_rs_b200 import FireflyCodec
from fish_speech_rs_b200._ffi import FsbError
from oracle import codec as ocodec

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PCM_ATOL = 1e-4


@pytest.fixture(scope="module")
def codec(codec_weights):
    c = FireflyCodec(codec_weights, max_frames=300, with_encoder=True)
    yield c
    c.close()


@pytest.mark.parametrize("T", [1, 2, 5, 33])
def test_decode_matches_oracle(codec, codec_weights, T):
    rng = np.random.default_rng(7 + T)
    codes = rng.integers(0, 1000, size=(1, 8, T)).astype(np.uint32)
    with torch.no_grad():
        exp = ocodec.decode(torch.from_numpy(codes.astype(np.int64)), codec_weights).numpy()
    got = codec.decode(codes)
    assert got.shape == (1, 1, 2048 * T)
    assert np.abs(exp).max() > 0.05 and np.abs(exp).max() < 0.999  # signal present, tanh not saturated
    err = np.abs(got - exp).max()
    assert err <= PCM_ATOL, f"max abs PCM error {err}"


def test_decode_default_voice_prefix(codec, codec_weights):
    """Real Fish-1.5 codes (voices-template/default.npy) as vocoder input."""
    v = np.load(os.path.join(GOLDEN, "default_voice.npy"))[:, :48]
    with torch.no_grad():
        exp = ocodec.decode(torch.from_numpy(v)[None], codec_weights).numpy()
    got = codec.decode(v[None].astype(np.uint32))
    assert np.abs(got - exp).max() <= PCM_ATOL


def test_decode_is_causal_and_chunk_consistent(codec):
    """Full-size property: decode(codes[:k]) == decode(codes)[:2048 k] (strictly causal stack)."""
    rng = np.random.default_rng(1)
    codes = rng.integers(0, 1000, size=(1, 8, 216)).astype(np.uint32)
    full = codec.decode(codes)
    part = codec.decode(codes[:, :, :100])
    np.testing.assert_array_equal(part, full[:, :, : 2048 * 100])


def test_decode_batch_equals_single(codec):
    rng = np.random.default_rng(2)
    cs = [rng.integers(0, 1000, size=(8, T)).astype(np.uint32) for T in (3, 9, 1)]
    outs = codec.decode_batch(cs)
    for c, o in zip(cs, outs):
        np.testing.assert_array_equal(o, codec.decode(c[None]))


def test_invalid_codes_rejected(codec):
    codes = np.full((1, 8, 4), 1000, np.uint32)  # Q11
    with pytest.raises(FsbError) as e:
        codec.decode(codes)
    assert e.value.status == -1
    with pytest.raises(FsbError):
        codec.decode(np.zeros((1, 8, 301), np.uint32))  # > max_frames


@pytest.mark.parametrize("Lm", [16, 37, 203])
def test_encode_mel_matches_oracle(codec, codec_weights, Lm):
    rng = np.random.default_rng(Lm)
    mel = (rng.standard_normal((1, 160, Lm)) * 2.0 - 4.0).astype(np.float32)  # log-mel-like range
    with torch.no_grad():
        exp = ocodec.encode_mel(torch.from_numpy(mel), codec_weights).numpy()
    got = codec.encode_mel(mel)
    assert got.shape == exp.shape == (1, 8, ((Lm - 2) // 2 + 1 - 2) // 2 + 1)
    # integer indices: exact except where a pre-round value sits within float noise of .5
    mism = (got != exp).mean()
    assert mism <= 0.002, f"{mism:.4f} of the indices differ"


@pytest.mark.parametrize("n", [6000, 512 * 40 - 1536, 44100 + 89])
def test_log_mel_matches_oracle(codec, n):
    """STFT (f64 DFT, periodic Hann, reflect pad) + mel table + clamp/log on the GPU == oracle.mel (numpy f64 rfft).
    Lengths: a short clip, an exact multiple of the hop (no partial chunk), one second + a partial chunk."""
    from oracle import mel as omel
    rng = np.random.default_rng(n)
    t = np.arange(n) / 44100.0
    pcm = (0.3 * np.sin(2 * np.pi * 220.0 * t) + 0.05 * rng.standard_normal(n)).astype(np.float32)
    pcm[n // 2: n // 2 + 700] = 0.0  # a silent stretch exercises the 1e-6 / clamp floor
    exp = omel.log_mel(pcm)
    got = codec.log_mel(pcm)
    assert got.shape == (1, 160, omel.n_mel_frames(n))
    # fp32 tolerance on the log-mel: 5e-4 absolute (values span [-11.5, 4.6]); measured 5e-7 on a 13 s clip
    np.testing.assert_allclose(got[0], exp, atol=5e-4, rtol=0)


def test_encode_from_pcm_equals_encode_of_the_oracle_mel(codec, codec_weights):
    """FireflyCodec::encode (firefly.rs:36-39): pcm -> codes with the mel kept on the device == encoder on the oracle's mel."""
    from oracle import mel as omel
    rng = np.random.default_rng(5)
    n = 30000
    pcm = (0.2 * rng.standard_normal(n)).astype(np.float32)
    m = omel.log_mel(pcm)
    with torch.no_grad():
        exp = ocodec.encode_mel(torch.from_numpy(m[None]), codec_weights).numpy()
    got = codec.encode(pcm)
    assert got.shape == exp.shape
    assert (got != exp).mean() <= 0.005


def test_block_decode_is_bit_identical_to_the_whole_utterance(codec, codec_weights):
    """Streaming: frames [t0, t1) decoded with a 16-frame causal halo == the same samples of decode(codes), bit for
    bit (the decoder's receptive field is 14.7 code frames); s16 + 44.1 -> 24 kHz resample on the device == the oracle
    output stage applied to the float PCM of that block."""
    from oracle import audio_out as ao
    rng = np.random.default_rng(3)
    T = 57
    codes = rng.integers(0, 1000, size=(1, 8, T)).astype(np.uint32)
    full = codec.decode(codes)[0, 0]
    for t0, t1 in ((0, 9), (9, 30), (30, 41), (41, 57), (17, 18)):
        blk = codec.decode_block(codes, t0, t1)[0, 0]
        np.testing.assert_array_equal(blk, full[2048 * t0: 2048 * t1])
        np.testing.assert_array_equal(codec.decode_block_s16(codes, t0, t1), ao.to_i16(blk))
        np.testing.assert_array_equal(codec.decode_block_s16(codes, t0, t1, to_rate=24000),
                                      ao.to_i16(ao.resample(blk, 44100, 24000)))


def _digits(idx):
    idx = np.asarray(idx)
    return np.stack([idx % 8, (idx // 8) % 5, (idx // 40) % 5, (idx // 200) % 5], -1)


def assert_index_mismatches_are_rounding_ties(got, pre, exp, tol=2e-3):
    """E3 bar (integer output): every index that differs from the oracle's must differ by exactly +-1 in digits whose
    pre-round value (the oracle's twice-bounded projection, fsq.rs:68-130) lies within `tol` of a rounding boundary x.5;
    all other digits must agree.  `pre`: (G, L, 4); got/exp: (1, G, L)."""
    bad = np.argwhere(got != exp)
    for _, g, l in bad:
        dg, de = _digits(got[0, g, l]), _digits(exp[0, g, l])
        for d in range(4):
            if dg[d] != de[d]:
                v = float(pre[g, l, d])
                frac = abs(v - np.floor(v) - 0.5)
                assert abs(int(dg[d]) - int(de[d])) == 1 and frac <= tol, \
                    f"group {g} frame {l} digit {d}: {dg[d]} vs {de[d]}, pre-round {v:.6f} is not a tie"
    return len(bad)


def test_fsq_encode_mismatches_are_proven_rounding_ties(codec, codec_weights):
    """Replaces the blanket mismatch budget of the tests above with a proof per mismatch."""
    total = 0
    for Lm in (37, 203, 611):
        rng = np.random.default_rng(100 + Lm)
        mel = (rng.standard_normal((1, 160, Lm)) * 2.0 - 4.0).astype(np.float32)
        with torch.no_grad():
            pre, exp = ocodec.encode_mel_preround(torch.from_numpy(mel), codec_weights)
        got = codec.encode_mel(mel)
        assert got.shape == tuple(exp.shape)
        n = assert_index_mismatches_are_rounding_ties(got, pre.numpy(), exp.numpy())
        assert n <= 0.005 * got.size
        total += got.size
    assert total > 1500


def test_sky_wav_through_the_gpu_front_end(codec, codec_weights):
    """tests/golden/sky.wav (the reference's tests/resources/sky.wav) through `FireflyCodec::encode` on the GPU: log-mel
    (160 x 1099) against oracle.mel, codes (1, 8, 274) against the oracle encoder on the oracle's mel."""
    import wave
    from oracle import mel as omel
    with wave.open(os.path.join(os.path.dirname(__file__), "golden", "sky.wav")) as w:
        pcm = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").astype(np.float32) / 32768.0
    exp_mel = omel.log_mel(pcm)
    got_mel = codec.log_mel(pcm)
    assert got_mel.shape == (1, 160, 1099)
    np.testing.assert_allclose(got_mel[0], exp_mel, atol=5e-4, rtol=0)
    with torch.no_grad():
        pre, exp = ocodec.encode_mel_preround(torch.from_numpy(exp_mel[None]), codec_weights)
    got = codec.encode(pcm)
    assert got.shape == (1, 8, 274)
    assert_index_mismatches_are_rounding_ties(got, pre.numpy(), exp.numpy())


def test_tensor_core_vocoder_agrees_with_the_fp32_fma_vocoder(codec_weights):
    """The HiFi-GAN ResBlock / upsampling convs run on tcgen05 (fp16 hi + lo split, three products per MAC,
    csrc/fsb_tc_conv.cu); FSB_CODEC_NO_TC=1 keeps the round-2a FP32-FMA kernels.  Both must sit inside the PCM
    tolerance of the oracle and close to each other, on a length that exercises several tiles per stage, a ragged
    last tile and the persistent tile loop (148 CTAs < tiles at the late stages)."""
    T = 97
    codes = np.random.default_rng(1234).integers(0, 1000, size=(1, 8, T)).astype(np.uint32)
    with torch.no_grad():
        exp = ocodec.decode(torch.from_numpy(codes.astype(np.int64)), codec_weights).numpy()
    outs, launches = {}, {}
    for name, env in (("tc", None), ("fma", "1")):
        old = os.environ.pop("FSB_CODEC_NO_TC", None)
        if env is not None:
            os.environ["FSB_CODEC_NO_TC"] = env
        try:
            c = FireflyCodec(codec_weights, max_frames=T)
            outs[name] = c.decode(codes)
            launches[name] = c.stats()["kernel_launches"]
            c.close()
        finally:
            os.environ.pop("FSB_CODEC_NO_TC", None)
            if old is not None:
                os.environ["FSB_CODEC_NO_TC"] = old
    for name, got in outs.items():
        err = np.abs(got - exp).max()
        assert err <= PCM_ATOL, f"{name}: max abs PCM error {err}"
    assert np.abs(outs["tc"] - outs["fma"]).max() <= 5e-5
    assert not np.array_equal(outs["tc"], outs["fma"])  # two different code paths really ran

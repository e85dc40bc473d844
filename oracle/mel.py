"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's log-mel front-end,
`LogMelSpectrogram::forward` (fish_speech_core/lib/audio/spectrogram.rs:29-83,141-158) on top of the streaming
STFT of audio/stft.rs:52-90 (rustfft f64, periodic Hann, overlap-and-save in hop-sized chunks).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this package.

Pinned against the reference's own data: `mel_filterbank()` reproduces `audio/melfilters160.bytes` (the table the
reference embeds with include_bytes!) to 1.8e-7 max-abs -- see tests/golden/make_mel_golden.py, which ran in the
build container where /root/reference is mounted, and tests/golden/mel_golden.json which it wrote.  The FFT itself is
numpy's f64 rfft (the reference: rustfft 6.2.0 in f64, un-vendored); both are exact to ~1e-13 relative."""
import numpy as np

N_FFT, HOP, N_MELS, SAMPLE_RATE = 2048, 512, 160, 44100
N_FREQS = N_FFT // 2 + 1


def _hz_to_mel(f):
    """Slaney scale (the table equals torchaudio melscale_fbanks(..., norm='slaney', mel_scale='slaney'))."""
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3.0
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, f / f_sp)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3.0
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(n_freqs=N_FREQS, n_mels=N_MELS, sample_rate=SAMPLE_RATE, f_min=0.0, f_max=None) -> np.ndarray:
    """(n_freqs, n_mels) f32 triangular filters with Slaney area normalisation (spectrogram.rs:85-96 loads it)."""
    f_max = float(sample_rate // 2) if f_max is None else f_max
    all_freqs = np.linspace(0.0, float(sample_rate // 2), n_freqs)
    f_pts = _mel_to_hz(np.linspace(_hz_to_mel(f_min), _hz_to_mel(f_max), n_mels + 2))
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    fb *= (2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels]))[None, :]
    return fb.astype(np.float32)


def reflect_pad(x: np.ndarray, pad: int) -> np.ndarray:
    """spectrogram.rs:14-27: the edge sample is repeated (numpy's 'symmetric', not 'reflect')."""
    return np.concatenate([x[:pad][::-1], x, x[len(x) - pad:][::-1]])


def n_mel_frames(n_samples: int) -> int:
    """Frames the streaming STFT emits: one per hop-sized chunk once 2048 samples have been pushed, plus one for a
    final partial chunk (stft.rs:52-90, spectrogram.rs:44-66)."""
    lp = n_samples + (N_FFT - HOP)
    full, rem = divmod(lp, HOP)
    return max(full - (N_FFT // HOP - 1), 0) + (1 if rem > 0 and lp >= N_FFT else 0)


def linear_spectrogram(x: np.ndarray) -> np.ndarray:
    """f32 (N) -> f32 (frames, 1025): |FFT| of the last 2048 samples after every chunk, + 1e-6."""
    xp = reflect_pad(np.asarray(x, np.float32), (N_FFT - HOP) // 2).astype(np.float64)
    nf = n_mel_frames(len(x))
    window = 0.5 * (1.0 - np.cos(2.0 * np.pi * np.arange(N_FFT) / N_FFT))
    frames = np.zeros((nf, N_FFT), np.float64)
    for f in range(nf):
        seg = xp[f * HOP: f * HOP + N_FFT]
        frames[f, : len(seg)] = seg  # a final partial chunk is zero-padded (stft.rs:61-64)
    spec = np.fft.rfft(frames * window[None, :], axis=1)
    mag = np.sqrt(spec.real ** 2 + spec.imag ** 2).astype(np.float32)
    return mag + np.float32(1e-6)


def log_mel(x: np.ndarray) -> np.ndarray:
    """LogMelSpectrogram::forward: f32 (N) -> f32 (160, frames)."""
    lin = linear_spectrogram(x)
    mel = lin @ mel_filterbank()  # f32 matmul, spectrogram.rs:131-136
    return np.log(np.clip(mel, np.float32(1e-5), np.float32(100.0))).T.astype(np.float32).copy()

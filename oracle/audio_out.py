"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's output stage --
`resample` (fish_speech_core/lib/audio/functional.rs:3-37), `Sample::to_i16` and `write_pcm_as_wav`
(fish_speech_core/lib/audio/wav.rs:9-13,27-58).  Only tests/, smoke() and bench.py's CPU legs may import it.
Pinned by closed-form known answers (tests/test_oracle_kat.py): the reference has no fixtures for these functions."""
import struct

import numpy as np


def resample(x: np.ndarray, from_rate: int, to_rate: int) -> np.ndarray:
    """f32 (n) -> f32 (ceil(n * to / from)): linear interpolation; indices in f64, weights in f32, v0*(1-t) + v1*t."""
    x = np.asarray(x, np.float32)
    n = len(x)
    ratio = np.float64(to_rate) / np.float64(from_rate)
    n_out = int(np.ceil(np.float64(n) * ratio))
    idx = np.arange(n_out, dtype=np.float64) / ratio
    i0 = np.floor(idx).astype(np.int64)
    i1 = np.minimum(np.ceil(idx).astype(np.int64), n - 1)
    t = (idx - np.floor(idx)).astype(np.float32)
    omt = (np.float32(1.0) - t).astype(np.float32)
    return (x[i0] * omt + x[i1] * t).astype(np.float32)


def to_i16(x: np.ndarray) -> np.ndarray:
    """(x.clamp(-1, 1) * 32767.0) as i16: truncation toward zero, NaN -> 0 (Rust `as`)."""
    x = np.asarray(x, np.float32)
    v = np.clip(np.where(np.isnan(x), np.float32(0), x), np.float32(-1), np.float32(1)) * np.float32(32767.0)
    return np.trunc(v).astype(np.int16)


def wav_bytes(samples_i16: np.ndarray, sample_rate: int) -> bytes:
    """write_pcm_as_wav: 44-byte RIFF header (mono, 16 bit) + little-endian samples."""
    n = len(samples_i16)
    total = 12 + 24 + n * 2 + 8
    hdr = b"RIFF" + struct.pack("<I", total - 8) + b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, sample_rate,
                                                                                 sample_rate * 2, 2, 16)
    return hdr + b"data" + struct.pack("<I", n * 2) + np.asarray(samples_i16, "<i2").tobytes()

"""Host-side mirror of `fish_speech_core::codec::FireflyCodec`
(fish_speech_core/lib/codec/firefly.rs:10-49) over the C ABI."""
import ctypes as C
from typing import Dict, List, Optional

import numpy as np

from . import _ffi as F


class FireflyCodec:
    """`FireflyCodec::load(cfg, vb, version)`; `.decode`, `.encode` (+ `.log_mel`, `.encode_mel`), `.sample_rate`."""

    def __init__(self, weights: Dict[str, "object"], fish_version: str = "1.5", device: int = 0,
                 max_frames: int = 512, with_encoder: bool = False, stream: int = 0):
        opts = F.fsb_codec_options(device, stream or None, {"1.4": F.FSB_FISH_1_4, "1.5": F.FSB_FISH_1_5}[fish_version],
                                   max_frames, int(with_encoder))
        table, keep = F.tensor_table(weights)
        h = C.c_void_p()
        F.check(F.lib().fsb_codec_create(table, len(weights), C.byref(opts), C.byref(h)))
        del keep
        self._h = h
        self.max_frames = max_frames
        self.sample_rate = F.lib().fsb_codec_sample_rate(h)

    def close(self):
        if getattr(self, "_h", None):
            F.lib().fsb_codec_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def decode(self, codes: np.ndarray) -> np.ndarray:
        """codes u32 (1, 8, T) -> pcm f32 (1, 1, 2048*T)  (firefly.rs:42-48)."""
        codes = np.ascontiguousarray(codes, dtype=np.uint32)
        if codes.ndim == 2:
            codes = codes[None]
        if codes.shape[0] != 1 or codes.shape[1] != 8:
            raise ValueError("codes must be (1, 8, T): the reference decode is only valid for batch 1")
        T = codes.shape[2]
        pcm = np.empty((1, 1, 2048 * T), np.float32)
        F.check(F.lib().fsb_codec_decode(self._h, codes.ctypes.data, T, pcm.ctypes.data))
        return pcm

    def decode_batch(self, codes: List[np.ndarray], out: Optional[List[np.ndarray]] = None) -> List[np.ndarray]:
        """n independent batch-1 decodes (Q9).  `out`: caller-owned (e.g. pinned) f32 buffers to fill."""
        cs = [np.ascontiguousarray(c, dtype=np.uint32).reshape(8, -1) for c in codes]
        n = len(cs)
        if out is None:
            outs = [np.empty((1, 1, 2048 * c.shape[1]), np.float32) for c in cs]
        else:
            outs = [o.reshape(-1)[: 2048 * c.shape[1]].reshape(1, 1, -1) for o, c in zip(out, cs)]
            assert all(o.dtype == np.float32 and o.flags.c_contiguous for o in outs)
        cp = (C.c_void_p * n)(*[c.ctypes.data for c in cs])
        op = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
        nf = (C.c_int32 * n)(*[c.shape[1] for c in cs])
        F.check(F.lib().fsb_codec_decode_batch(self._h, cp, nf, n, op))
        return outs

    def encode_mel(self, mel: np.ndarray) -> np.ndarray:
        """log-mel f32 (1, 160, Lm) -> codes i64 (1, 8, L)  (encoder.rs:38-42)."""
        mel = np.ascontiguousarray(mel, dtype=np.float32)
        if mel.ndim == 2:
            mel = mel[None]
        Lm = mel.shape[2]
        cap = Lm // 4 + 1
        out = np.zeros((1, 8, cap), np.int64)
        n = C.c_size_t()
        F.check(F.lib().fsb_codec_encode_mel(self._h, mel.ctypes.data, Lm, out.ctypes.data, cap, C.byref(n)))
        return out[:, :, : n.value].copy()

    def decode_block(self, codes: np.ndarray, t0: int, t1: int) -> np.ndarray:
        """frames [t0, t1) of codes u32 (1, 8, T) -> pcm f32 (1, 1, 2048*(t1-t0)), identical to the same samples of
        `decode(codes)` (causal decoder, 16-frame left halo): the streaming path of speech.rs:180-236."""
        codes = np.ascontiguousarray(np.asarray(codes, dtype=np.uint32).reshape(8, -1))
        out = np.zeros((1, 1, 2048 * (t1 - t0)), np.float32)
        F.check(F.lib().fsb_codec_decode_block(self._h, codes.ctypes.data, codes.shape[1], t0, t1, out.ctypes.data))
        return out

    def decode_block_s16(self, codes: np.ndarray, t0: int, t1: int, to_rate: int = 0) -> np.ndarray:
        """same block as s16 for the wire, optionally resampled on the device (functional.rs:3-37, wav.rs:9-13)."""
        codes = np.ascontiguousarray(np.asarray(codes, dtype=np.uint32).reshape(8, -1))
        cap = 2048 * (t1 - t0) + 16
        out = np.zeros(cap, np.int16)
        n = C.c_size_t()
        F.check(F.lib().fsb_codec_decode_block_s16(self._h, codes.ctypes.data, codes.shape[1], t0, t1, to_rate,
                                                   out.ctypes.data, cap, C.byref(n)))
        return out[: n.value].copy()

    def log_mel(self, pcm: np.ndarray) -> np.ndarray:
        """mono 44.1 kHz pcm f32 (n) -> log-mel f32 (1, 160, Lm)  (`LogMelSpectrogram::forward`, spectrogram.rs:141-158)."""
        pcm = np.ascontiguousarray(pcm, dtype=np.float32).reshape(-1)
        cap = pcm.size // 512 + 8
        out = np.zeros((160, cap), np.float32)
        n = C.c_size_t()
        F.check(F.lib().fsb_codec_log_mel(self._h, pcm.ctypes.data, pcm.size, out.ctypes.data, cap, C.byref(n)))
        # the library writes (160, Lm) densely
        return out.reshape(-1)[: 160 * n.value].reshape(1, 160, n.value).copy()

    def encode(self, pcm: np.ndarray) -> np.ndarray:
        """`FireflyCodec::encode` (firefly.rs:36-39): pcm f32 (1, 1, n) -> codes i64 (1, 8, L); the mel stays on the device."""
        pcm = np.ascontiguousarray(pcm, dtype=np.float32).reshape(-1)
        cap = pcm.size // 2048 + 8
        out = np.zeros((1, 8, cap), np.int64)
        n = C.c_size_t()
        F.check(F.lib().fsb_codec_encode(self._h, pcm.ctypes.data, pcm.size, out.ctypes.data, cap, C.byref(n)))
        return out[:, :, : n.value].copy()

    def stats(self) -> Dict:
        s = F.fsb_codec_stats()
        F.check(F.lib().fsb_codec_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in s._fields_}

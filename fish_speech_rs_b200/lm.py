"""Host-side mirror of `fish_speech_core::lm` over the C ABI (include/fsb.h).

Names, argument meaning and error behaviour follow the reference:
  DualARTransformer        fish_speech_core/lib/lm/dual_ar.rs:443-713
  SamplingArgs             fish_speech_core/lib/lm/sampling/mod.rs:28-34
  generate_blocking        fish_speech_core/lib/lm/generate/single_batch.rs:308-324
  generate_static_batch    fish_speech_core/lib/lm/generate/static_batch.rs:282-390
All compute happens in libfsb.so on the GPU; nothing here has a CPU path.
"""
import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _ffi as F


@dataclass
class SamplingArgs:
    temp: float = 0.7
    top_p: float = 0.8
    top_k: int = 256
    repetition_penalty: float = 1.4
    seed: int = 0  # Philox key; the reference seeds StdRng from entropy (single_batch.rs:46) / 42 (static_batch.rs:63)

    def _c(self):
        return F.fsb_sampling_args(float(self.temp), float(self.top_p), int(self.top_k),
                                   float(self.repetition_penalty), int(self.seed))


_VERSIONS = {"1.2": F.FSB_FISH_1_2, "1.4": F.FSB_FISH_1_4, "1.5": F.FSB_FISH_1_5}


class DualARTransformer:
    """`DualARTransformer::load(vb, cfg, token_config, model_type)` + methods."""

    def __init__(self, weights: Dict[str, "object"], cfg: Dict, token_config: Dict, fish_version: str = "1.5",
                 device: int = 0, dtype: str = "f32", max_batch: int = 1, max_seq_len: int = 0, stream: int = 0,
                 decode_mode: int = 0):
        """decode_mode: 0 auto, 1 per-op kernels replayed from a CUDA graph, 2 persistent megakernel."""
        self.cfg = dict(cfg)
        self.token_config = dict(token_config)
        self.model_type = fish_version
        args = F.fsb_model_args(
            int(bool(cfg.get("attention_qkv_bias", False))), cfg["codebook_size"], cfg["dim"], cfg["head_dim"],
            int(cfg.get("intermediate_size") or 0), cfg["max_seq_len"], cfg["n_fast_layer"], cfg["n_head"],
            cfg["n_layer"], cfg["n_local_heads"], cfg["num_codebooks"], cfg["vocab_size"],
            int(bool(cfg.get("tie_word_embeddings", False))), float(cfg["norm_eps"]), float(cfg["rope_base"]))
        end = token_config.get("semantic_end_id")
        tok = F.fsb_token_config(token_config["im_end_id"], token_config["pad_id"], token_config["semantic_start_id"],
                                 0 if end is None else end, 0 if end is None else 1)
        opts = F.fsb_lm_options(device, stream or None, {"f32": F.FSB_F32, "bf16": F.FSB_BF16}[dtype], max_batch,
                                max_seq_len, _VERSIONS[fish_version], decode_mode)
        table, keep = F.tensor_table(weights)
        h = C.c_void_p()
        F.check(F.lib().fsb_lm_create(C.byref(args), C.byref(tok), table, len(weights), C.byref(opts), C.byref(h)))
        del keep
        self._h = h
        self.max_batch = max_batch
        self.dtype = dtype

    def close(self):
        if getattr(self, "_h", None):
            F.lib().fsb_lm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- step API (dual_ar.rs:574-673)
    def forward_generate(self, inp: np.ndarray, input_pos: int, want_logits: bool = True
                         ) -> Tuple[Optional[np.ndarray], np.ndarray]:
        """inp u32 (B, C+1, S) -> (logits (B,1,V), hidden (B,1,D) pre-norm)."""
        inp = np.ascontiguousarray(inp, dtype=np.uint32)
        if inp.ndim == 2:
            inp = inp[None]
        B, C1, S = inp.shape
        if C1 != self.cfg["num_codebooks"] + 1:
            raise ValueError(f"expected {self.cfg['num_codebooks'] + 1} rows, got {C1}")
        logits = np.empty((B, 1, self.cfg["vocab_size"]), np.float32) if want_logits else None
        hidden = np.empty((B, 1, self.cfg["dim"]), np.float32)
        F.check(F.lib().fsb_lm_forward_generate(self._h, inp.ctypes.data, B, S, input_pos,
                                                logits.ctypes.data if want_logits else None, hidden.ctypes.data))
        return logits, hidden

    def forward_generate_fast(self, x: np.ndarray, input_pos: int) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, 1, self.cfg["dim"])
        out = np.empty((x.shape[0], 1, self.cfg["codebook_size"]), np.float32)
        F.check(F.lib().fsb_lm_forward_generate_fast(self._h, x.ctypes.data, x.shape[0], input_pos, out.ctypes.data))
        return out

    def fast_embeddings(self, ids: Sequence[int]) -> np.ndarray:
        ids = np.ascontiguousarray(ids, dtype=np.uint32).reshape(-1)
        out = np.empty((ids.shape[0], self.cfg["dim"]), np.float32)
        F.check(F.lib().fsb_lm_fast_embeddings(self._h, ids.ctypes.data, ids.shape[0], out.ctypes.data))
        return out

    def clear_fast_layer_caches(self):
        F.check(F.lib().fsb_lm_clear_fast_layer_caches(self._h))

    def clear_slow_layer_caches(self):
        F.check(F.lib().fsb_lm_clear_slow_layer_caches(self._h))

    def clear_slow_caches_until(self, pos: int):
        F.check(F.lib().fsb_lm_clear_slow_caches_until(self._h, pos))

    def curr_kv_size(self) -> int:
        n = C.c_size_t()
        F.check(F.lib().fsb_lm_curr_kv_size(self._h, C.byref(n)))
        return n.value

    def last_frames(self, row: int = 0, cap: int = 0) -> np.ndarray:
        """Frames of the last generate call for `row` as `SingleBatchGenerator::next` yields them: u32 (C+1, T),
        slow token in row 0, <|im_end|> frames included (single_batch.rs:76-214)."""
        cap = cap or (self.cfg["max_seq_len"] + 2)
        out = np.zeros((self.cfg["num_codebooks"] + 1, cap), np.uint32)
        n = C.c_size_t()
        F.check(F.lib().fsb_lm_last_frames(self._h, int(row), out.ctypes.data, cap, C.byref(n)))
        return out[:, : n.value].copy()

    # ---- per-voice conditioning KV (speech.rs:40)
    def kv_snapshot_save(self, row: int, n_positions: int):
        h = C.c_void_p()
        F.check(F.lib().fsb_lm_kv_snapshot_save(self._h, int(row), int(n_positions), C.byref(h)))
        return h

    def kv_snapshot_restore(self, snap, row: int):
        F.check(F.lib().fsb_lm_kv_snapshot_restore(self._h, snap, int(row)))

    def kv_snapshot_free(self, snap):
        F.check(F.lib().fsb_lm_kv_snapshot_free(self._h, snap))

    # ---- continuous batching (state.rs:13, static_batch.rs:160-173)
    def session_begin(self, sampling_args: "SamplingArgs", fixed_len: bool = False):
        sa = sampling_args._c()
        F.check(F.lib().fsb_lm_session_begin(self._h, C.byref(sa), F.FSB_GEN_FIXED_LEN if fixed_len else 0))

    def session_admit(self, slot: int, prompt: np.ndarray, max_new_tokens: int, fixed_len: int = 0):
        prompt = np.ascontiguousarray(prompt, dtype=np.uint32)
        F.check(F.lib().fsb_lm_session_admit(self._h, int(slot), prompt.ctypes.data, prompt.shape[1], int(max_new_tokens),
                                             int(fixed_len)))

    def session_run(self, max_frames: int) -> np.ndarray:
        """up to `max_frames` more frames for every live slot; returns the per-slot "still generating" mask"""
        act = (C.c_int32 * self.max_batch)()
        n = C.c_int32()
        F.check(F.lib().fsb_lm_session_run(self._h, int(max_frames), act, C.byref(n)))
        return np.array(list(act), dtype=bool)

    def session_collect(self, slot: int, cap: int) -> np.ndarray:
        Cb = self.cfg["num_codebooks"]
        out = np.zeros((Cb, cap), np.uint32)
        n = C.c_size_t()
        F.check(F.lib().fsb_lm_session_collect(self._h, int(slot), out.ctypes.data, cap, C.byref(n)))
        return out[:, : n.value].copy()

    def set_profile(self, on: bool):
        F.check(F.lib().fsb_lm_set_profile(self._h, int(on)))

    def stats(self) -> Dict:
        s = F.fsb_lm_stats()
        F.check(F.lib().fsb_lm_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in s._fields_}


def generate_blocking(model: DualARTransformer, prompt: np.ndarray, max_new_tokens: int, sampling_args: SamplingArgs,
                      fixed_len: Optional[int] = None, keep_slow_kv: bool = False) -> np.ndarray:
    """prompt u32 (C+1, P) -> codes u32 (C, T).  `fixed_len` is the bench harness flag."""
    prompt = np.ascontiguousarray(prompt, dtype=np.uint32)
    Cb = model.cfg["num_codebooks"]
    if prompt.ndim != 2 or prompt.shape[0] != Cb + 1:
        raise ValueError(f"prompt must be ({Cb + 1}, P)")
    P = prompt.shape[1]
    cap = max(int(max_new_tokens) - P + 2, 1) if fixed_len is None else int(fixed_len)
    out = np.zeros((Cb, cap), np.uint32)
    n = C.c_size_t()
    flags = (F.FSB_GEN_FIXED_LEN if fixed_len is not None else 0) | (F.FSB_GEN_KEEP_SLOW_KV if keep_slow_kv else 0)
    sa = sampling_args._c()
    F.check(F.lib().fsb_lm_generate_blocking(model._h, prompt.ctypes.data, P, int(max_new_tokens), C.byref(sa), flags,
                                             int(fixed_len or 0), out.ctypes.data, cap, C.byref(n)))
    return out[:, : n.value].copy()


def generate_blocking_with_hidden(model: DualARTransformer, prompt: np.ndarray, max_new_tokens: int,
                                  sampling_args: SamplingArgs, fixed_len: Optional[int] = None
                                  ) -> Tuple[np.ndarray, np.ndarray]:
    """`generate_blocking_with_hidden(..., collect_hidden_states=true)` (single_batch.rs:217-306): codes u32 (C, T) and
    the pre-norm slow hidden state of every yielded frame, f32 (T_all, 1, dim)."""
    prompt = np.ascontiguousarray(prompt, dtype=np.uint32)
    Cb, D = model.cfg["num_codebooks"], model.cfg["dim"]
    P = prompt.shape[1]
    cap = max(int(max_new_tokens) - P + 2, 1) if fixed_len is None else int(fixed_len)
    out = np.zeros((Cb, cap), np.uint32)
    hid = np.zeros((cap + 1, D), np.float32)
    n, nh = C.c_size_t(), C.c_size_t()
    flags = F.FSB_GEN_FIXED_LEN if fixed_len is not None else 0
    sa = sampling_args._c()
    F.check(F.lib().fsb_lm_generate_blocking_with_hidden(model._h, prompt.ctypes.data, P, int(max_new_tokens), C.byref(sa),
                                                         flags, int(fixed_len or 0), out.ctypes.data, cap, C.byref(n),
                                                         hid.ctypes.data, cap + 1, C.byref(nh)))
    return out[:, : n.value].copy(), hid[: nh.value].reshape(nh.value, 1, D).copy()


def generate_static_batch(model: DualARTransformer, prompts: List[np.ndarray], max_new_tokens: int,
                          sampling_args: SamplingArgs, fixed_len: Optional[int] = None) -> List[np.ndarray]:
    """Row i == generate_blocking(prompts[i]) with Philox row index i (independent utterances)."""
    B = len(prompts)
    Cb = model.cfg["num_codebooks"]
    ps = [np.ascontiguousarray(p, dtype=np.uint32) for p in prompts]
    for p in ps:
        if p.ndim != 2 or p.shape[0] != Cb + 1:
            raise ValueError(f"every prompt must be ({Cb + 1}, P)")
    lens = (C.c_int32 * B)(*[p.shape[1] for p in ps])
    minP = min(p.shape[1] for p in ps)
    cap = max(int(max_new_tokens) - minP + 2, 1) if fixed_len is None else int(fixed_len)
    outs = [np.zeros((Cb, cap), np.uint32) for _ in range(B)]
    pp = (C.c_void_p * B)(*[p.ctypes.data for p in ps])
    op = (C.c_void_p * B)(*[o.ctypes.data for o in outs])
    out_lens = (C.c_size_t * B)()
    flags = F.FSB_GEN_FIXED_LEN if fixed_len is not None else 0
    sa = sampling_args._c()
    F.check(F.lib().fsb_lm_generate_static_batch(model._h, pp, lens, B, int(max_new_tokens), C.byref(sa), flags,
                                                 int(fixed_len or 0), op, cap, out_lens))
    return [o[:, : out_lens[i]].copy() for i, o in enumerate(outs)]

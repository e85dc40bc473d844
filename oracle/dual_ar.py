"""Dual-AR transformer -- oracle (test infrastructure only; PyTorch CPU fp32).

Op-for-op restatement of fish_speech_core/lib/lm/dual_ar.rs, *as written*:
growing `cat` KV cache (:316-324), materialised x8 K/V repeat (:327-357), scale
folded into K^T (:260), full-vocabulary output head (:629-634), pre-norm hidden
state handed to the fast stack (Q1).  Candle op semantics restated from the
published candle-nn / candle-core 0.8.3 sources (un-vendored, "recalled" in
SURVEY.md section 8c): rope_i interleaved pairs, rms_norm with f32 accumulation,
softmax_last_dim max-subtracted, silu = x / (1 + exp(-x)).
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class BaseModelArgs:
    """dual_ar.rs:56-116 (serde names kept)."""
    attention_qkv_bias: bool = False
    codebook_size: int = 1024
    dim: int = 1024
    dropout: float = 0.0
    head_dim: int = 64
    initializer_range: float = 0.02
    intermediate_size: Optional[int] = 4096
    max_seq_len: int = 8192
    model_type: str = "dual_ar"
    n_fast_layer: int = 4
    n_head: int = 16
    n_layer: int = 24
    n_local_heads: int = 2
    norm_eps: float = 1e-6
    num_codebooks: int = 8
    rope_base: float = 1000000.0
    tie_word_embeddings: bool = False
    use_gradient_checkpointing: bool = True
    vocab_size: int = 102048
    depthwise_wte: Optional[bool] = None
    depthwise_output: Optional[bool] = None

    @property
    def inter(self) -> int:
        return self.intermediate_size or self.dim * 4


@dataclass
class TokenConfig:
    """dual_ar.rs:17-54.  Fish 1.5: semantic_end_id is Some; <=1.4: None and
    semantic_start_id is the single <|semantic|> id."""
    im_end_id: int
    pad_id: int
    semantic_start_id: int
    semantic_end_id: Optional[int]


def precompute_freqs_cis(cfg: BaseModelArgs) -> Tuple[torch.Tensor, torch.Tensor]:
    """dual_ar.rs:168-186.  theta_i = 1 / base^(i/n) (f32 powf), table = pos * theta
    (f32 product), then f32 cos / sin.  Transcendentals are evaluated in f64 and
    rounded once to f32 (== correctly rounded f32 libm), the same recipe the CUDA
    host code uses so both tables are bit-identical."""
    n_elem = cfg.dim // cfg.n_head
    i = np.arange(0, n_elem, 2, dtype=np.float32)
    expo = (i / np.float32(n_elem)).astype(np.float32)
    powv = np.power(np.float64(np.float32(cfg.rope_base)), expo.astype(np.float64)).astype(np.float32)
    theta = (np.float32(1.0) / powv).astype(np.float32)
    pos = np.arange(cfg.max_seq_len, dtype=np.float32)
    idx_theta = (pos[:, None] * theta[None, :]).astype(np.float32)
    cos = np.cos(idx_theta.astype(np.float64)).astype(np.float32)
    sin = np.sin(idx_theta.astype(np.float64)).astype(np.float32)
    return torch.from_numpy(cos), torch.from_numpy(sin)


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """candle_nn::RmsNorm: x / sqrt(mean(x^2) + eps) * w."""
    ms = x.pow(2).mean(dim=-1, keepdim=True)
    return x / torch.sqrt(ms + eps) * w


def rope_i(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """candle_nn::rotary_emb::rope_i -- interleaved pairs.  x (B,H,S,hd), cos/sin (S,hd/2)."""
    x0 = x[..., 0::2]
    x1 = x[..., 1::2]
    o0 = x0 * cos - x1 * sin
    o1 = x0 * sin + x1 * cos
    return torch.stack([o0, o1], dim=-1).flatten(-2)


def silu(x: torch.Tensor) -> torch.Tensor:
    return x / (1.0 + torch.exp(-x))


def get_mask_abs(size1: int, size2: int, context: int) -> torch.Tensor:
    """dual_ar.rs:702-712.  1 == MASK."""
    i = torch.arange(size1)[:, None]
    j = torch.arange(size2)[None, :]
    return ((size1 + j > size2 + i) | (size1 + j + context < size2 + i)).to(torch.uint8)


class Attention:
    """dual_ar.rs:197-405."""

    def __init__(self, w: Dict[str, torch.Tensor], prefix: str, cfg: BaseModelArgs):
        self.cfg = cfg
        self.wqkv = w[prefix + "wqkv.weight"]
        self.wo = w[prefix + "wo.weight"]
        self.kv_cache: Optional[Tuple[torch.Tensor, torch.Tensor]] = None

    def forward(self, x, mask, cos, sin):
        cfg = self.cfg
        bsz, seqlen, _ = x.shape
        H, KV, hd = cfg.n_head, cfg.n_local_heads, cfg.head_dim
        qkv = F.linear(x, self.wqkv)
        q = qkv[..., : H * hd].reshape(bsz, seqlen, H, hd).transpose(1, 2)
        k = qkv[..., H * hd: (H + KV) * hd].reshape(bsz, seqlen, KV, hd).transpose(1, 2)
        v = qkv[..., (H + KV) * hd:].reshape(bsz, seqlen, KV, hd).transpose(1, 2)
        q = rope_i(q.contiguous(), cos, sin)
        k = rope_i(k.contiguous(), cos, sin)
        v = v.contiguous()
        if self.kv_cache is not None:
            k = torch.cat([self.kv_cache[0], k], dim=2)
            v = torch.cat([self.kv_cache[1], v], dim=2)
        self.kv_cache = (k, v)
        kv_len = k.shape[2]
        n_rep = H // KV
        k_rep = k.unsqueeze(2).expand(bsz, KV, n_rep, kv_len, hd).reshape(bsz, H, kv_len, hd)
        v_rep = v.unsqueeze(2).expand(bsz, KV, n_rep, kv_len, hd).reshape(bsz, H, kv_len, hd)
        scale = np.float32(1.0) / np.sqrt(np.float32(hd))
        att = q @ (k_rep.transpose(-1, -2) * float(scale))
        if seqlen > 1:
            att = att.masked_fill(mask.bool(), float("-inf"))
        att = torch.softmax(att, dim=-1)
        y = att @ v_rep.contiguous()
        y = y.transpose(1, 2).reshape(bsz, seqlen, H * hd)
        return F.linear(y, self.wo)

    def clear_cache(self):
        self.kv_cache = None

    def clear_cache_until(self, pos: int):
        if self.kv_cache is not None:
            k, v = self.kv_cache
            n = min(k.shape[2], pos)
            self.kv_cache = (k[:, :, :n].contiguous(), v[:, :, :n].contiguous())


class TransformerBlock:
    """dual_ar.rs:407-441."""

    def __init__(self, w, prefix, cfg):
        self.cfg = cfg
        self.attention = Attention(w, prefix + "attention.", cfg)
        self.w1 = w[prefix + "feed_forward.w1.weight"]
        self.w2 = w[prefix + "feed_forward.w2.weight"]
        self.w3 = w[prefix + "feed_forward.w3.weight"]
        self.ffn_norm = w[prefix + "ffn_norm.weight"]
        self.attention_norm = w[prefix + "attention_norm.weight"]

    def forward(self, x, mask, cos, sin):
        eps = self.cfg.norm_eps
        h = x + self.attention.forward(rms_norm(x, self.attention_norm, eps), mask, cos, sin)
        n = rms_norm(h, self.ffn_norm, eps)
        ff = F.linear(silu(F.linear(n, self.w1)) * F.linear(n, self.w3), self.w2)
        return h + ff


class DualARTransformer:
    """dual_ar.rs:443-713."""

    def __init__(self, weights: Dict[str, torch.Tensor], cfg: BaseModelArgs,
                 token_config: TokenConfig, fish_version: str = "1.5"):
        self.cfg = cfg
        self.token_config = token_config
        self.model_type = fish_version
        w = {k: v.to(torch.float32) for k, v in weights.items()}
        self.embeddings = w["embeddings.weight"]
        self.codebook_embeddings = w["codebook_embeddings.weight"]
        self.layers = [TransformerBlock(w, f"layers.{l}.", cfg) for l in range(cfg.n_layer)]
        self.norm = w["norm.weight"]
        self.output = w["embeddings.weight" if cfg.tie_word_embeddings else "output.weight"]
        self.fast_embeddings = w["fast_embeddings.weight"]
        self.fast_layers = [TransformerBlock(w, f"fast_layers.{l}.", cfg) for l in range(cfg.n_fast_layer)]
        self.fast_norm = w["fast_norm.weight"]
        self.fast_output = w["fast_output.weight"]
        self.freqs_cis = precompute_freqs_cis(cfg)

    def embed(self, x: torch.Tensor) -> torch.Tensor:
        """dual_ar.rs:532-567.  x (B, C+1, S) integer."""
        cfg, tc = self.cfg, self.token_config
        assert x.shape[-2] == cfg.num_codebooks + 1
        sem = x[:, 0, :].long()
        codes = x[:, 1:, :].long()
        out = self.embeddings[sem]  # (B,S,D)
        if tc.semantic_end_id is not None:
            m = (sem <= tc.semantic_end_id) & (sem >= tc.semantic_start_id)
        else:
            m = sem == tc.semantic_start_id
        m = m.to(torch.float32).unsqueeze(-1)
        for c in range(cfg.num_codebooks):
            out = out + self.codebook_embeddings[codes[:, c, :] + c * cfg.codebook_size] * m
        return out

    def curr_kv_size(self) -> int:
        kv = self.layers[0].attention.kv_cache
        return 0 if kv is None else kv[0].shape[-2]

    def forward_generate(self, inp: torch.Tensor, input_pos: int):
        """dual_ar.rs:574-635 -> (logits (B,1,V), hidden (B,1,D) pre-norm)."""
        x = self.embed(inp)
        _, seq_len, _ = x.shape
        if seq_len == 1:
            mask = get_mask_abs(1, 1, self.cfg.max_seq_len)
        else:
            mask = get_mask_abs(seq_len, self.curr_kv_size() + seq_len, self.cfg.max_seq_len)
        cos, sin = self.freqs_cis
        c, s = cos[input_pos: input_pos + seq_len], sin[input_pos: input_pos + seq_len]
        for layer in self.layers:
            x = layer.forward(x, mask, c, s)
        x = x[:, seq_len - 1: seq_len, :]
        logits = F.linear(rms_norm(x, self.norm, self.cfg.norm_eps), self.output)
        return logits, x

    def forward_generate_fast(self, x: torch.Tensor, input_pos: int) -> torch.Tensor:
        """dual_ar.rs:638-673 -> (B,1,codebook_size)."""
        cos, sin = self.freqs_cis
        seq_len = x.shape[1]
        c, s = cos[input_pos: input_pos + seq_len], sin[input_pos: input_pos + seq_len]
        for layer in self.fast_layers:
            x = layer.forward(x, None, c, s)
        return F.linear(rms_norm(x, self.fast_norm, self.cfg.norm_eps), self.fast_output)

    def clear_fast_layer_caches(self):
        for l in self.fast_layers:
            l.attention.clear_cache()

    def clear_slow_layer_caches(self):
        for l in self.layers:
            l.attention.clear_cache()

    def clear_slow_caches_until(self, pos: int):
        for l in self.layers:
            l.attention.clear_cache_until(pos)

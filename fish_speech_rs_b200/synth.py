"""Seeded synthetic checkpoints and prompts (no weights exist offline).

Weight names and shapes follow the reference loaders
(dual_ar.rs:125-156,219-223,415-419,466-511; SURVEY.md Appendix A for the codec).
Everything is generated on the CPU with torch's seeded generator so the oracle,
the tests and bench.py all see the same tensors.
"""
import math
import re
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch

FISH15 = dict(attention_qkv_bias=False, codebook_size=1024, dim=1024, dropout=0.0, head_dim=64,
              initializer_range=0.02, intermediate_size=4096, max_seq_len=8192, model_type="dual_ar",
              n_fast_layer=4, n_head=16, n_layer=24, n_local_heads=2, norm_eps=1e-6, num_codebooks=8,
              rope_base=1000000.0, tie_word_embeddings=False, use_gradient_checkpointing=True,
              vocab_size=102048)
FISH14 = dict(FISH15, vocab_size=32000, max_seq_len=4096)
# Token ids: Fish 1.5 tokenizer has <|im_end|> = 100011 directly before <|semantic:0|> = 100012
# (dual_ar.rs:36 fallback), semantic range 1024 wide.
FISH15_TOKENS = dict(im_end_id=100011, pad_id=5, semantic_start_id=100012, semantic_end_id=100012 + 1023)
FISH14_TOKENS = dict(im_end_id=4, pad_id=5, semantic_start_id=5, semantic_end_id=None)

# A small config with the same structure, for tests the CPU oracle finishes in seconds.
TINY = dict(FISH15, dim=256, n_layer=3, n_fast_layer=2, n_head=4, n_local_heads=2, head_dim=64,
            intermediate_size=512, vocab_size=2304, max_seq_len=512, codebook_size=1024)
# Full-width blocks (dim 1024, FFN 4096, 16/2 heads) with few layers and a small vocabulary: the shapes the
# single-row decode megakernel (TMA weight ring) is specialised for, still cheap enough for the CPU oracle.
WIDE = dict(FISH15, n_layer=2, n_fast_layer=2, vocab_size=2304, max_seq_len=512)
TINY_TOKENS = dict(im_end_id=1263, pad_id=1200, semantic_start_id=1264, semantic_end_id=1264 + 1023)


def lm_weight_shapes(cfg: Dict) -> Dict[str, tuple]:
    D, I = cfg["dim"], cfg["intermediate_size"] or cfg["dim"] * 4
    H, KV, hd = cfg["n_head"], cfg["n_local_heads"], cfg["head_dim"]
    V, C, CS = cfg["vocab_size"], cfg["num_codebooks"], cfg["codebook_size"]
    shapes = {"embeddings.weight": (V, D), "codebook_embeddings.weight": (C * CS, D)}

    def block(p):
        shapes[p + "attention.wqkv.weight"] = ((H + 2 * KV) * hd, D)
        shapes[p + "attention.wo.weight"] = (D, D)
        shapes[p + "feed_forward.w1.weight"] = (I, D)
        shapes[p + "feed_forward.w2.weight"] = (D, I)
        shapes[p + "feed_forward.w3.weight"] = (I, D)
        shapes[p + "ffn_norm.weight"] = (D,)
        shapes[p + "attention_norm.weight"] = (D,)

    for l in range(cfg["n_layer"]):
        block(f"layers.{l}.")
    shapes["norm.weight"] = (D,)
    if not cfg["tie_word_embeddings"]:
        shapes["output.weight"] = (V, D)
    shapes["fast_embeddings.weight"] = (CS, D)
    for l in range(cfg["n_fast_layer"]):
        block(f"fast_layers.{l}.")
    shapes["fast_norm.weight"] = (D,)
    shapes["fast_output.weight"] = (CS, D)
    return shapes


def make_lm_weights(cfg: Dict, seed: int = 1234, round_bf16: bool = False) -> Dict[str, torch.Tensor]:
    """normal(0, initializer_range) matrices; norm weights 1 + 0.05 * normal."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in lm_weight_shapes(cfg).items():
        if len(shape) == 1:
            t = 1.0 + 0.05 * torch.randn(shape, generator=g)
        else:
            t = torch.randn(shape, generator=g) * cfg["initializer_range"]
        if round_bf16:
            t = t.to(torch.bfloat16).to(torch.float32)
        out[name] = t.contiguous()
    return out


def codec_weight_shapes(with_encoder: bool = True) -> Dict[str, tuple]:
    s: Dict[str, tuple] = {}

    def convnext(p, dim):
        s[p + "dwconv.conv.weight"] = (dim, 1, 7)
        s[p + "dwconv.conv.bias"] = (dim,)
        s[p + "norm.weight"] = (dim,)
        s[p + "norm.bias"] = (dim,)
        s[p + "pwconv1.weight"] = (4 * dim, dim)
        s[p + "pwconv1.bias"] = (4 * dim,)
        s[p + "pwconv2.weight"] = (dim, 4 * dim)
        s[p + "pwconv2.bias"] = (dim,)
        s[p + "gamma"] = (dim,)

    for g in range(8):
        p = f"quantizer.residual_fsq.rvqs.{g}."
        s[p + "project_in.weight"] = (4, 64)
        s[p + "project_in.bias"] = (4,)
        s[p + "project_out.weight"] = (64, 4)
        s[p + "project_out.bias"] = (64,)
    for i in range(2):
        s[f"quantizer.downsample.{i}.0.conv.weight"] = (512, 512, 2)
        s[f"quantizer.downsample.{i}.0.conv.bias"] = (512,)
        convnext(f"quantizer.downsample.{i}.1.", 512)
        s[f"quantizer.upsample.{i}.0.conv.weight"] = (512, 512, 2)
        s[f"quantizer.upsample.{i}.0.conv.bias"] = (512,)
        convnext(f"quantizer.upsample.{i}.1.", 512)
    s["head.conv_pre.conv.weight"] = (512, 512, 13)
    s["head.conv_pre.conv.bias"] = (512,)
    ks = (16, 16, 4, 4, 4)
    for i in range(5):
        cin, cout = 512 // 2 ** i, 512 // 2 ** (i + 1)
        s[f"head.ups.{i}.conv.weight"] = (cin, cout, ks[i])
        s[f"head.ups.{i}.conv.bias"] = (cout,)
        for j, k in enumerate((3, 7, 11)):
            for m in range(3):
                for cv in ("convs1", "convs2"):
                    s[f"head.resblocks.{i}.blocks.{j}.{cv}.{m}.conv.weight"] = (cout, cout, k)
                    s[f"head.resblocks.{i}.blocks.{j}.{cv}.{m}.conv.bias"] = (cout,)
    s["head.conv_post.conv.weight"] = (1, 16, 13)
    s["head.conv_post.conv.bias"] = (1,)
    if with_encoder:
        dims, depths = (128, 256, 384, 512), (3, 3, 9, 3)
        s["backbone.downsample_layers.0.0.conv.weight"] = (128, 160, 7)
        s["backbone.downsample_layers.0.0.conv.bias"] = (128,)
        s["backbone.downsample_layers.0.1.weight"] = (128,)
        s["backbone.downsample_layers.0.1.bias"] = (128,)
        for i in range(1, 4):
            s[f"backbone.downsample_layers.{i}.0.weight"] = (dims[i - 1],)
            s[f"backbone.downsample_layers.{i}.0.bias"] = (dims[i - 1],)
            s[f"backbone.downsample_layers.{i}.1.weight"] = (dims[i], dims[i - 1], 1)
            s[f"backbone.downsample_layers.{i}.1.bias"] = (dims[i],)
        for i in range(4):
            for j in range(depths[i]):
                convnext(f"backbone.stages.{i}.{j}.", dims[i])
        s["backbone.norm.weight"] = (512,)
        s["backbone.norm.bias"] = (512,)
    return s


def make_codec_weights(seed: int = 4321, with_encoder: bool = True) -> Dict[str, torch.Tensor]:
    """normal(0, gain / sqrt(fan_in)) convs/linears, small biases, gamma ~ 0.1,
    LayerNorm weight ~ 1.  Gains keep the pre-tanh PCM at O(1) so tanh does not
    saturate and hide errors."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in codec_weight_shapes(with_encoder).items():
        if name.endswith("gamma"):
            t = 0.1 + 0.02 * torch.randn(shape, generator=g)
        elif name.endswith("norm.weight") or (len(shape) == 1 and name.endswith(".weight")):
            t = 1.0 + 0.05 * torch.randn(shape, generator=g)
        elif name.endswith("bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        else:
            if name.startswith("head.ups.") or re.match(r"quantizer\.upsample\.\d\.0\.conv\.weight", name):
                # ConvTranspose (in, out, k): k/stride taps per output
                stride = {16: 8, 4: 2, 2: 2}[shape[2]]
                fan_in = shape[0] * shape[2] // stride
            elif len(shape) == 3:
                fan_in = shape[1] * shape[2]
            else:
                fan_in = shape[1]
            gain = 1.0 if "resblocks" in name else (2.0 if name.startswith("head.ups.") else (0.25 if "conv_post" in name else 1.0))
            t = torch.randn(shape, generator=g) * (gain / math.sqrt(fan_in))
        out[name] = t.contiguous()
    return out


def default_voice_codes(n: int = 274, seed: int = 99) -> np.ndarray:
    """Stand-in for voices-template/default.npy (int64 (8,274), values 3..999); the
    real file lives in the read-only reference tree and is committed as a fixture
    under tests/golden/default_voice.npy."""
    rng = np.random.default_rng(seed)
    return rng.integers(3, 1000, size=(8, n), dtype=np.int64)


def make_prompt(cfg: Dict, tok: Dict, P: int, seed: int, voice: Optional[np.ndarray] = None) -> np.ndarray:
    """SURVEY.md 8d synthetic prompt, layout of text/prompt.rs:33-105: u32 (C+1, P):
    [text span | VQ span (row0 = semantic_start + codes[0], rows 1..C = codes) | 4 text columns].
    Text ids are drawn below im_end_id so no text column is mistaken for a semantic token."""
    C = cfg["num_codebooks"]
    rng = np.random.default_rng(seed)
    if voice is None:
        voice = default_voice_codes()
    voice = voice[:C]
    nv = min(voice.shape[1], max(P - 8, 0))
    prompt = np.zeros((C + 1, P), dtype=np.uint32)
    text_hi = min(tok["im_end_id"], 100000)
    n_text = P - nv - 4
    prompt[0, :n_text] = rng.integers(0, text_hi, size=n_text)
    if nv > 0:
        if tok.get("semantic_end_id") is not None:
            prompt[0, n_text:n_text + nv] = tok["semantic_start_id"] + voice[0, :nv]
            prompt[1:, n_text:n_text + nv] = voice[:, :nv]
        else:  # Fish <= 1.4: single <|semantic|> id, codes stored +1 (prompt.rs:83-90)
            prompt[0, n_text:n_text + nv] = tok["semantic_start_id"]
            prompt[1:, n_text:n_text + nv] = voice[:, :nv] + 1
    prompt[0, P - 4:] = rng.integers(0, text_hi, size=4)
    return prompt

"""CPU test: the oracle's transformer stacks against an INDEPENDENT implementation of the same published architecture.

The reference cannot run here (Rust + Candle, no toolchain) and holds no golden activations, so the arithmetic of
oracle/dual_ar.py is otherwise checked only against closed forms.  Both stacks of the dual-AR model are Llama blocks
(RMSNorm -> GQA attention with RoPE -> SwiGLU, dual_ar.rs:125-165,232-341,414-443); Hugging Face `transformers` ships
a Llama written by other people.  The only difference is the RoPE pairing: the reference rotates interleaved pairs
(x[2j], x[2j+1]) (`rope_i`, dual_ar.rs:246-247), HF rotates (x[j], x[j + d/2]); permuting the rows of wq / wk inside
each head maps one onto the other without changing q.k.  With that permutation the two must agree to fp32 rounding:
prefill, an incremental step on the KV cache (RoPE offset + cache append + mask), and the fast stack fed one position
at a time."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from fish_speech_rs_b200 import synth
from oracle import dual_ar as olm

transformers = pytest.importorskip("transformers")


def hf_llama(cfg, w, prefix, n_layers, norm_key, permute=True):
    from transformers import LlamaConfig, LlamaModel
    D, hd, H, KV = cfg["dim"], cfg["head_dim"], cfg["n_head"], cfg["n_local_heads"]
    kw = dict(vocab_size=8, hidden_size=D, intermediate_size=cfg["intermediate_size"], num_hidden_layers=n_layers,
              num_attention_heads=H, num_key_value_heads=KV, head_dim=hd, rms_norm_eps=cfg["norm_eps"],
              attention_bias=False, mlp_bias=False, max_position_embeddings=cfg["max_seq_len"], hidden_act="silu",
              attn_implementation="eager")
    try:
        conf = LlamaConfig(rope_theta=cfg["rope_base"], **kw)
    except TypeError:  # newer releases keep the base inside rope_parameters
        conf = LlamaConfig(rope_parameters={"rope_type": "default", "rope_theta": cfg["rope_base"]}, **kw)
    m = LlamaModel(conf).to(torch.float32).eval()
    # interleaved pair (2j, 2j+1) of a head  ->  HF's (j, j + hd/2)
    perm = torch.tensor([2 * r if r < hd // 2 else 2 * (r - hd // 2) + 1 for r in range(hd)])

    def heads(rows, n):
        return rows.reshape(n, hd, D)[:, perm, :].reshape(n * hd, D) if permute else rows

    sd = {}
    for l in range(n_layers):
        wqkv = torch.as_tensor(np.asarray(w[f"{prefix}.{l}.attention.wqkv.weight"], dtype=np.float32))
        q, k, v = wqkv[: H * hd], wqkv[H * hd: (H + KV) * hd], wqkv[(H + KV) * hd:]
        t = lambda name: torch.as_tensor(np.asarray(w[f"{prefix}.{l}.{name}"], dtype=np.float32))  # noqa: E731
        sd[f"layers.{l}.self_attn.q_proj.weight"] = heads(q, H)
        sd[f"layers.{l}.self_attn.k_proj.weight"] = heads(k, KV)
        sd[f"layers.{l}.self_attn.v_proj.weight"] = v
        sd[f"layers.{l}.self_attn.o_proj.weight"] = t("attention.wo.weight")
        sd[f"layers.{l}.mlp.gate_proj.weight"] = t("feed_forward.w1.weight")
        sd[f"layers.{l}.mlp.up_proj.weight"] = t("feed_forward.w3.weight")
        sd[f"layers.{l}.mlp.down_proj.weight"] = t("feed_forward.w2.weight")
        sd[f"layers.{l}.input_layernorm.weight"] = t("attention_norm.weight")
        sd[f"layers.{l}.post_attention_layernorm.weight"] = t("ffn_norm.weight")
    sd["norm.weight"] = torch.as_tensor(np.asarray(w[norm_key], dtype=np.float32))
    sd["embed_tokens.weight"] = torch.zeros(8, D)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all("rotary" in k or "inv_freq" in k for k in missing), (missing, unexpected)
    return m


@pytest.fixture(scope="module", params=["TINY", "WIDE"])  # WIDE: full-width Fish blocks (dim 1024, 16 q / 2 kv heads, FFN 4096)
def setup(request):
    cfg, tok = dict(getattr(synth, request.param)), dict(synth.TINY_TOKENS)
    w = synth.make_lm_weights(cfg, seed=1234)
    ora = olm.DualARTransformer(w, olm.BaseModelArgs(**cfg), olm.TokenConfig(**tok))
    return cfg, tok, w, ora


def test_slow_stack_prefill_and_cached_step_match_hf_llama(setup):
    cfg, tok, w, ora = setup
    hf = hf_llama(cfg, w, "layers", cfg["n_layer"], "norm.weight")
    P = 41
    prompt = torch.from_numpy(synth.make_prompt(cfg, tok, P + 1, seed=7).astype(np.int64))[None]  # (1, C+1, P+1)
    with torch.no_grad():
        ora.clear_slow_layer_caches()
        emb = ora.embed(prompt)                                               # (1, P+1, D): the stack's input
        _, h_pre = ora.forward_generate(prompt[:, :, :P], 0)                  # prefill of P positions
        _, h_step = ora.forward_generate(prompt[:, :, P:], P)                 # one step on the KV cache
        ref = hf(inputs_embeds=emb).last_hidden_state                         # (1, P+1, D), final norm applied
        got_pre = olm.rms_norm(h_pre, ora.norm, cfg["norm_eps"])
        got_step = olm.rms_norm(h_step, ora.norm, cfg["norm_eps"])
    ora.clear_slow_layer_caches()
    scale = float(ref.abs().max())
    assert scale > 0.1
    np.testing.assert_allclose(got_pre[0, 0].numpy(), ref[0, P - 1].numpy(), atol=5e-6 * max(scale, 1.0), rtol=0)
    np.testing.assert_allclose(got_step[0, 0].numpy(), ref[0, P].numpy(), atol=5e-6 * max(scale, 1.0), rtol=0)
    # negative control: without the pair permutation the same weights give a different function (the check has teeth)
    with torch.no_grad():
        wrong = hf_llama(cfg, w, "layers", cfg["n_layer"], "norm.weight", permute=False)(inputs_embeds=emb).last_hidden_state
    assert float((got_pre[0, 0] - wrong[0, P - 1]).abs().max()) > 1e-2


def test_fast_stack_steps_match_hf_llama(setup):
    cfg, tok, w, ora = setup
    hf = hf_llama(cfg, w, "fast_layers", cfg["n_fast_layer"], "fast_norm.weight")
    g = torch.Generator().manual_seed(3)
    xs = torch.randn(1, 5, cfg["dim"], generator=g)                           # five codebook positions
    with torch.no_grad():
        ora.clear_fast_layer_caches()
        got = [ora.forward_generate_fast(xs[:, i: i + 1], i) for i in range(5)]  # (1, 1, codebook_size) each
        ref = F.linear(hf(inputs_embeds=xs).last_hidden_state, ora.fast_output)
    ora.clear_fast_layer_caches()
    scale = max(float(ref.abs().max()), 1.0)
    for i in range(5):
        np.testing.assert_allclose(got[i][0, 0].numpy(), ref[0, i].numpy(), atol=5e-6 * scale, rtol=0)

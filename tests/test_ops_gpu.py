"""Operator-level entry points (the CustomOp1 precedent, lm/ops/repeat_kv.rs)."""
import ctypes as C

import numpy as np
import pytest
import torch

from fish_speech_rs_b200 import _ffi as F
from oracle import dual_ar as olm

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seqlen", [1, 166])  # the reference's own test shapes, repeat_kv.rs:126-162
def test_repeat_kv_matches_cat_reference(seqlen):
    x = torch.randn(1, 2, seqlen, 64, device="cuda")
    out = torch.empty(1, 16, seqlen, 64, device="cuda")
    F.check(F.lib().fsb_op_repeat_kv(x.data_ptr(), out.data_ptr(), F.FSB_F32, 2, 8, seqlen, 64,
                                     torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    exp = torch.cat([x[:, h:h + 1].expand(1, 8, seqlen, 64) for h in range(2)], dim=1)
    assert torch.equal(out, exp)
    xb = x.to(torch.bfloat16)
    ob = torch.empty(1, 16, seqlen, 64, device="cuda", dtype=torch.bfloat16)
    F.check(F.lib().fsb_op_repeat_kv(xb.data_ptr(), ob.data_ptr(), F.FSB_BF16, 2, 8, seqlen, 64,
                                     torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert torch.equal(ob, exp.to(torch.bfloat16))


@pytest.mark.parametrize("kv_len", [0, 1, 165, 599, 1699])
def test_gqa_decode_attn_matches_oracle(kv_len):
    """rope_i + cat + repeat_kv + matmul + softmax + matmul (dual_ar.rs:239-249,316-376) in one call."""
    B, H, KV, hd, max_len = 2, 16, 2, 64, 2048
    g = torch.Generator().manual_seed(kv_len)
    qkv = torch.randn(B, (H + 2 * KV) * hd, generator=g)
    kc = torch.zeros(B, KV, max_len, hd)
    vc = torch.zeros(B, KV, max_len, hd)
    kc[:, :, :kv_len] = torch.randn(B, KV, kv_len, hd, generator=g)
    vc[:, :, :kv_len] = torch.randn(B, KV, kv_len, hd, generator=g)
    cos, sin = olm.precompute_freqs_cis(olm.BaseModelArgs(max_seq_len=max_len))
    # oracle
    q = qkv[:, : H * hd].reshape(B, 1, H, hd).transpose(1, 2)
    k = qkv[:, H * hd:(H + KV) * hd].reshape(B, 1, KV, hd).transpose(1, 2)
    v = qkv[:, (H + KV) * hd:].reshape(B, 1, KV, hd).transpose(1, 2)
    q = olm.rope_i(q, cos[kv_len:kv_len + 1], sin[kv_len:kv_len + 1])
    k = olm.rope_i(k, cos[kv_len:kv_len + 1], sin[kv_len:kv_len + 1])
    kk = torch.cat([kc[:, :, :kv_len], k], dim=2).repeat_interleave(H // KV, dim=1)
    vv = torch.cat([vc[:, :, :kv_len], v], dim=2).repeat_interleave(H // KV, dim=1)
    att = torch.softmax(q @ (kk.transpose(-1, -2) * 0.125), dim=-1) @ vv
    exp = att.transpose(1, 2).reshape(B, H * hd)
    # device
    d = lambda t: t.contiguous().cuda()
    qkv_d, kc_d, vc_d, cos_d, sin_d = d(qkv), d(kc), d(vc), d(cos), d(sin)
    pos = torch.full((B,), kv_len, dtype=torch.int32, device="cuda")
    out = torch.empty(B, H * hd, device="cuda")
    nbytes = F.lib().fsb_op_gqa_decode_attn_scratch_bytes(B, H, hd)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    F.check(F.lib().fsb_op_gqa_decode_attn(qkv_d.data_ptr(), kc_d.data_ptr(), vc_d.data_ptr(), cos_d.data_ptr(),
                                           sin_d.data_ptr(), pos.data_ptr(), B, H, KV, hd, max_len, out.data_ptr(),
                                           scratch.data_ptr(), nbytes, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    np.testing.assert_allclose(out.cpu().numpy(), exp.numpy(), atol=2e-5, rtol=0)
    # the new K/V row was appended in place (Tensor::cat replaced)
    np.testing.assert_allclose(kc_d[:, :, kv_len].cpu().numpy(), k[:, :, 0].numpy(), atol=1e-6, rtol=0)
    np.testing.assert_array_equal(vc_d[:, :, kv_len].cpu().numpy(), v[:, :, 0].numpy())

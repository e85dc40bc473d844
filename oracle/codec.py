"""Firefly-GAN-VQ codec -- oracle (test infrastructure only; PyTorch CPU fp32).

Restates, for Fish >= 1.4 (causal convs), batch 1:
  codec/utils/mod.rs:53-63,110-122   FishConvNet / FishTransConvNet
  codec/convnext.rs:109-154,224-334  ConvNeXtBlock, LayerNormChannelsFirst, encoder
  codec/fsq.rs:68-159                FSQ bound / quantize / indices<->codes
  codec/grouped_residual_fsq.rs      ResidualFSQ / GroupedResidualFSQ
  codec/quantizer.rs:104-146         DownsampleFiniteScalarQuantizer
  codec/hifi_gan.rs:73-216           ResBlock1 / ParallelBlock / HiFiGAN
  codec/decoder.rs:37-68, encoder.rs:38-42, firefly.rs:36-48
Candle semantics restated (un-vendored, "recalled"): gelu() == tanh approximation
(Q10), LayerNorm biased variance, mean == sum * (1/n), round half away from zero.
"""
import math
from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

LEVELS = (8, 5, 5, 5)
UPSAMPLE_RATES = (8, 8, 2, 2, 2)
UPSAMPLE_KERNELS = (16, 16, 4, 4, 4)
RES_KERNELS = (3, 7, 11)
RES_DILATIONS = (1, 3, 5)
N_GROUPS = 8
DIM = 512
HOP = 512
ENC_DIMS = (128, 256, 384, 512)
ENC_DEPTHS = (3, 3, 9, 3)


def silu(x):
    return x / (1.0 + torch.exp(-x))


def gelu_tanh(x):
    return F.gelu(x, approximate="tanh")


def fish_conv(x, w, b, stride=1, dilation=1, groups=1):
    """FishConvNet::forward (>=1.4): left zero-pad (k-1)*d + 1 - stride, conv, no other padding."""
    k = w.shape[-1]
    pad = (k - 1) * dilation + 1 - stride
    x = F.pad(x, (pad, 0))
    return F.conv1d(x, w, b, stride=stride, dilation=dilation, groups=groups)


def fish_trans_conv(x, w, b, stride):
    """FishTransConvNet::forward (>=1.4): conv_transpose1d then drop k - stride samples from the right."""
    k = w.shape[-1]
    y = F.conv_transpose1d(x, w, b, stride=stride)
    trim = max(k - stride, 0)
    return y[..., : y.shape[-1] - trim] if trim else y


def convnext_block(x, w: Dict[str, torch.Tensor], p: str):
    """ConvNeXtBlock::forward, convnext.rs:109-127."""
    C = x.shape[1]
    h = fish_conv(x, w[p + "dwconv.conv.weight"], w[p + "dwconv.conv.bias"], groups=C)
    h = h.permute(0, 2, 1)
    h = F.layer_norm(h, (C,), w[p + "norm.weight"], w[p + "norm.bias"], eps=1e-6)
    h = gelu_tanh(F.linear(h, w[p + "pwconv1.weight"], w[p + "pwconv1.bias"]))
    h = F.linear(h, w[p + "pwconv2.weight"], w[p + "pwconv2.bias"])
    if p + "gamma" in w:
        h = w[p + "gamma"] * h
    return x + h.permute(0, 2, 1)


def layer_norm_channels_first(x, weight, bias, eps=1e-6):
    """LayerNormChannelsFirst::forward, convnext.rs:144-154."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    xn = (x - u) / torch.sqrt(s + eps)
    return xn * weight[:, None] + bias[:, None]


# ---------------------------------------------------------------- FSQ

def fsq_implicit_codebook() -> torch.Tensor:
    """FSQ::implicit_codebook (fsq.rs:154-159) in the reference's f32 arithmetic -> (1000, 4)."""
    levels = torch.tensor(LEVELS, dtype=torch.float32)
    basis = torch.tensor([1.0, 8.0, 40.0, 200.0], dtype=torch.float32)
    idx = torch.arange(0, 1000, dtype=torch.float32).unsqueeze(-1)
    cnc = torch.floor(idx / basis)
    cnc = cnc - torch.floor(cnc / levels) * levels
    half_width = torch.floor(levels / 2.0)
    return (cnc - half_width) / half_width


def fsq_bound(z):
    """FSQ::bound, fsq.rs:68-84 (eps = 1e-3)."""
    levels = torch.tensor(LEVELS, dtype=torch.float32)
    half_l = (levels - 1.0) * 1.001 / 2.0
    offset = ((levels - torch.floor(levels / 2.0) * 2.0) == 0).to(torch.float32) * 0.5
    r = offset / half_l
    shift = torch.log((1.0 + r) / (1.0 - r)) * 0.5
    return torch.tanh(z + shift) * half_l - offset


def round_half_away(x):
    return torch.sign(x) * torch.floor(torch.abs(x) + 0.5)


def fsq_encode_group(x, w, p):
    """ResidualFSQ::forward with one quantizer (grouped_residual_fsq.rs:75-93):
    project_in -> bound (implicit first step) -> FSQ::forward (bound again, round) -> index."""
    levels = torch.tensor(LEVELS, dtype=torch.float32)
    basis = torch.tensor([1.0, 8.0, 40.0, 200.0], dtype=torch.float32)
    half_width = torch.floor(levels / 2.0)
    z = F.linear(x, w[p + "project_in.weight"], w[p + "project_in.bias"])
    residual = fsq_bound(z)
    codes = round_half_away(fsq_bound(residual / 1.0)) / half_width
    zhat = codes * half_width + half_width
    return (zhat * basis).sum(-1).to(torch.int64)


def quantizer_decode(codes: torch.Tensor, w, prefix="quantizer.") -> torch.Tensor:
    """DownsampleFiniteScalarQuantizer::decode, quantizer.rs:126-146.  codes (1, 8, T)."""
    assert codes.shape[0] == 1, "reference decode is only valid for batch 1 (Q9)"
    cb = fsq_implicit_codebook()
    outs = []
    for g in range(N_GROUPS):
        idx = codes[0, g].long()
        if int(idx.max()) >= cb.shape[0]:
            raise IndexError("code >= 1000 is invalid for the FSQ table (Q11)")
        c = cb[idx]  # (T, 4)
        p = f"{prefix}residual_fsq.rvqs.{g}."
        outs.append(F.linear(c, w[p + "project_out.weight"], w[p + "project_out.bias"]))
    z = torch.cat(outs, dim=-1).unsqueeze(0).transpose(1, 2)  # (1, 512, T)
    for i in range(2):  # upsample.0 then upsample.1 (quantizer.rs:126-133)
        p = f"{prefix}upsample.{i}."
        z = fish_trans_conv(z, w[p + "0.conv.weight"], w[p + "0.conv.bias"], 2)
        z = convnext_block(z, w, p + "1.")
    return z


def quantizer_encode(z: torch.Tensor, w, prefix="quantizer.") -> torch.Tensor:
    """DownsampleFiniteScalarQuantizer::encode, quantizer.rs:104-124 -> (1, 8, L) int64."""
    for i in range(2):
        p = f"{prefix}downsample.{i}."
        z = fish_conv(z, w[p + "0.conv.weight"], w[p + "0.conv.bias"], stride=2)
        z = convnext_block(z, w, p + "1.")
    zt = z.transpose(1, 2)  # (1, L, 512)
    chunks = zt.chunk(N_GROUPS, dim=-1)
    idx = [fsq_encode_group(c, w, f"{prefix}residual_fsq.rvqs.{g}.") for g, c in enumerate(chunks)]
    return torch.stack(idx, dim=1)  # (1, 8, L)


# ---------------------------------------------------------------- HiFi-GAN head

def resblock1(x, w, p, k):
    """ResBlock1::forward (hifi_gan.rs:73-86); both convs dilated for >= 1.4 (:56-59)."""
    for m, d in enumerate(RES_DILATIONS):
        xt = silu(x)
        xt = silu(fish_conv(xt, w[f"{p}convs1.{m}.conv.weight"], w[f"{p}convs1.{m}.conv.bias"], dilation=d))
        xt = fish_conv(xt, w[f"{p}convs2.{m}.conv.weight"], w[f"{p}convs2.{m}.conv.bias"], dilation=d)
        x = x + xt
    return x


def hifigan_forward(x, w, prefix="head.", return_pre_tanh=False):
    """HiFiGAN::forward, hifi_gan.rs:207-216."""
    x = fish_conv(x, w[prefix + "conv_pre.conv.weight"], w[prefix + "conv_pre.conv.bias"])
    for i, (u, _k) in enumerate(zip(UPSAMPLE_RATES, UPSAMPLE_KERNELS)):
        x = fish_trans_conv(silu(x), w[f"{prefix}ups.{i}.conv.weight"], w[f"{prefix}ups.{i}.conv.bias"], u)
        rs = [resblock1(x, w, f"{prefix}resblocks.{i}.blocks.{j}.", k) for j, k in enumerate(RES_KERNELS)]
        x = (rs[0] + rs[1] + rs[2]) * float(np.float32(1.0 / 3.0))  # stack + mean(0)
    x = fish_conv(silu(x), w[prefix + "conv_post.conv.weight"], w[prefix + "conv_post.conv.bias"])
    return x if return_pre_tanh else torch.tanh(x)


def decode(codes: torch.Tensor, w: Dict[str, torch.Tensor]) -> torch.Tensor:
    """FireflyCodec::decode (firefly.rs:42-48) -> f32 (1, 1, 2048*T).  The length
    masks of FireflyDecoder::decode (decoder.rs:45-66) are all ones for b == 1."""
    z = quantizer_decode(codes, w)
    return hifigan_forward(z, w)


# ---------------------------------------------------------------- encoder (cfg4)

def convnext_encoder(mel, w, prefix="backbone."):
    """ConvNeXtEncoder::forward, convnext.rs:224-334.  mel (1, 160, Lm) -> (1, 512, Lm)."""
    p = prefix + "downsample_layers."
    x = fish_conv(mel, w[p + "0.0.conv.weight"], w[p + "0.0.conv.bias"])
    x = layer_norm_channels_first(x, w[p + "0.1.weight"], w[p + "0.1.bias"])
    for j in range(ENC_DEPTHS[0]):
        x = convnext_block(x, w, f"{prefix}stages.0.{j}.")
    for i in range(1, 4):
        x = layer_norm_channels_first(x, w[f"{p}{i}.0.weight"], w[f"{p}{i}.0.bias"])
        x = F.conv1d(x, w[f"{p}{i}.1.weight"], w[f"{p}{i}.1.bias"])
        for j in range(ENC_DEPTHS[i]):
            x = convnext_block(x, w, f"{prefix}stages.{i}.{j}.")
    return layer_norm_channels_first(x, w[prefix + "norm.weight"], w[prefix + "norm.bias"])


def encode_mel(mel, w):
    """FireflyEncoder::encode, encoder.rs:38-42."""
    return quantizer_encode(convnext_encoder(mel, w), w)


def encode_mel_preround(mel, w):
    """Same chain as `encode_mel`, but returns what FSQ rounds: the twice-bounded projections (G, L, 4) right before
    `round()` (fsq.rs:68-130), next to the indices.  Lets a test prove that an index that differs on another
    implementation differs only because a pre-round value sits within float noise of a rounding boundary (x.5)."""
    z = convnext_encoder(mel, w)
    prefix = "quantizer."
    for i in range(2):
        z = fish_conv(z, w[f"{prefix}downsample.{i}.0.conv.weight"], w[f"{prefix}downsample.{i}.0.conv.bias"], stride=2)
        z = convnext_block(z, w, f"{prefix}downsample.{i}.1.")
    zt = z[0].transpose(0, 1)  # (L, 512)
    pre, idx = [], []
    for g in range(N_GROUPS):
        p = f"{prefix}residual_fsq.rvqs.{g}."
        x = zt[:, g * 64:(g + 1) * 64]
        zz = F.linear(x, w[p + "project_in.weight"], w[p + "project_in.bias"])
        pre.append(fsq_bound(fsq_bound(zz)))
        idx.append(fsq_encode_group(x, w, p))
    return torch.stack(pre), torch.stack(idx)[None]

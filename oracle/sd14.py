"""fp32 PyTorch restatement of the Stable-Diffusion-1.4 arithmetic MADM reaches through
diffusers==0.25.0 (UNet2DConditionModel, AutoencoderKL.encoder, DDPMScheduler.add_noise).

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED by the reference.

Module / parameter names follow diffusers so ``state_dict`` keys are identical to the keys
MADM checkpoints hold under ``backbone.feature_extractor.ldm_extractor.{unet,vae}.*``
(reference ``checkpoint/odise_checkpointer.py:38-111``).  Control flow follows the reference's
re-implementation of the forward passes:

* ``modeling/meta_arch/ldm_diffusers.py:283-311``  vae_encoder (tap counter, deterministic mean)
* ``modeling/meta_arch/ldm_diffusers.py:349-360``  add_noise (shared noise)
* ``modeling/meta_arch/ldm_diffusers.py:454-616``  diffusion_unet (time path, down/mid/up, taps)
* ``modeling/meta_arch/ldm_diffusers.py:363-451``  up-block bodies with skip concat + taps

``Q``/``QR``/``QI`` are optional storage-rounding hooks (identity by default).  Tests set them to a
bf16 round-trip to budget the error of a bf16-storage pipeline against the fp32 oracle; with the
default identity hooks they do not change the oracle's results.
"""
import math
from typing import Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

_IDENT = lambda t: t  # noqa: E731
# storage-rounding emulation hooks; identity = exact fp32 oracle
Q = _IDENT    # tensors that are GEMM/conv operands (bf16 by construction in a bf16 tensor-core pipeline)
QR = _IDENT   # the residual stream (block outputs, skip connections)
QI = _IDENT   # intermediate GEMM outputs that only feed a normalisation


def set_storage_rounding(operand=None, residual=None, intermediate=None):
    """Install (or clear) the storage-rounding hooks used for error budgeting."""
    global Q, QR, QI
    Q = operand if operand is not None else _IDENT
    QR = residual if residual is not None else _IDENT
    QI = intermediate if intermediate is not None else _IDENT


def q(t):
    return Q(t)


def qr(t):
    return QR(t)


def qi(t):
    return QI(t)


# --------------------------------------------------------------------------------------
# building blocks (diffusers ResnetBlock2D / Attention / BasicTransformerBlock / ...)
# --------------------------------------------------------------------------------------
class ResnetBlock2D(nn.Module):
    """h = conv1(silu(GN32(x))); h += time_emb_proj(silu(temb)); h = conv2(silu(GN32(h)));
    out = conv_shortcut(x) + h   (dropout 0, output_scale_factor 1)."""

    def __init__(self, cin: int, cout: int, temb_channels: Optional[int], eps: float, groups: int = 32):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, cout) if temb_channels is not None else None
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps, affine=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb=None):
        h = q(F.silu(self.norm1(x)))
        h = self.conv1(h)
        if self.time_emb_proj is not None:
            h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = qi(h)
        h = q(F.silu(self.norm2(h)))
        h = self.conv2(h)
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(q(x))
        return qr(x + h)


class Attention(nn.Module):
    """diffusers Attention (AttnProcessor2_0 semantics): softmax(QK^T / sqrt(d)) V, no mask."""

    def __init__(self, query_dim: int, heads: int, cross_dim: Optional[int] = None, qkv_bias: bool = False,
                 group_norm: Optional[Tuple[int, float]] = None, residual: bool = False):
        super().__init__()
        kv_dim = cross_dim if cross_dim is not None else query_dim
        self.heads = heads
        self.residual = residual
        self.group_norm = nn.GroupNorm(group_norm[0], query_dim, eps=group_norm[1]) if group_norm else None
        self.to_q = nn.Linear(query_dim, query_dim, bias=qkv_bias)
        self.to_k = nn.Linear(kv_dim, query_dim, bias=qkv_bias)
        self.to_v = nn.Linear(kv_dim, query_dim, bias=qkv_bias)
        self.to_out = nn.ModuleList([nn.Linear(query_dim, query_dim), nn.Dropout(0.0)])

    def forward(self, x, ctx=None):
        res = x
        is4d = x.dim() == 4
        if is4d:  # VAE mid-block attention takes [B,C,H,W]
            b, c, hh, ww = x.shape
            if self.group_norm is not None:
                x = self.group_norm(x)
            x = x.view(b, c, hh * ww).transpose(1, 2)
        x = q(x)
        ctx = x if ctx is None else ctx
        b, n, c = x.shape
        d = c // self.heads
        qh = q(self.to_q(x)).view(b, n, self.heads, d).transpose(1, 2)
        kh = q(self.to_k(ctx)).view(b, -1, self.heads, d).transpose(1, 2)
        vh = q(self.to_v(ctx)).view(b, -1, self.heads, d).transpose(1, 2)
        o = F.scaled_dot_product_attention(qh, kh, vh)
        o = q(o.transpose(1, 2).reshape(b, n, c))
        o = self.to_out[0](o)
        if is4d:
            o = o.transpose(1, 2).reshape(b, c, hh, ww)
        if self.residual:
            o = qr(o + res)
        return o


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        h, g = self.proj(x).chunk(2, dim=-1)
        return q(h * F.gelu(g))  # exact (erf) GELU


class FeedForward(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Dropout(0.0), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, heads: int, cross_dim: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-5)
        self.attn2 = Attention(dim, heads, cross_dim=cross_dim)
        self.norm3 = nn.LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim)

    def forward(self, x, ctx):
        x = qr(x + self.attn1(q(self.norm1(x))))
        x = qr(x + self.attn2(q(self.norm2(x)), ctx))
        x = qr(x + self.ff(q(self.norm3(x))))
        return x


class Transformer2DModel(nn.Module):
    """GN32(eps 1e-6) -> conv1x1 proj_in -> [B,HW,C] -> BasicTransformerBlock -> conv1x1 proj_out -> + residual."""

    def __init__(self, dim: int, heads: int, cross_dim: int, groups: int = 32):
        super().__init__()
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(dim, dim, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(dim, heads, cross_dim)])
        self.proj_out = nn.Conv2d(dim, dim, 1)

    def forward(self, x, ctx):
        b, c, h, w = x.shape
        res = x
        x = q(self.norm(x))
        x = qr(self.proj_in(x))
        x = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
        x = self.transformer_blocks[0](x, ctx)
        x = x.reshape(b, h, w, c).permute(0, 3, 1, 2)
        return qr(self.proj_out(q(x)) + res)


class Downsample2D(nn.Module):
    def __init__(self, ch: int, padding: int):
        super().__init__()
        self.padding = padding
        self.conv = nn.Conv2d(ch, ch, 3, stride=2, padding=padding)

    def forward(self, x):
        if self.padding == 0:  # VAE encoder: asymmetric pad right/bottom, stride-2 pad-0 conv
            x = F.pad(x, (0, 1, 0, 1))
        return qr(self.conv(q(x)))


class Upsample2D(nn.Module):
    def __init__(self, ch: int):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return qr(self.conv(F.interpolate(q(x), scale_factor=2.0, mode="nearest")))


class TimestepEmbedding(nn.Module):
    def __init__(self, cin: int, dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(cin, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


def timestep_sinusoid(t: torch.Tensor, dim: int = 320, max_period: float = 10000.0) -> torch.Tensor:
    """diffusers Timesteps(320, flip_sin_to_cos=True, downscale_freq_shift=0): [cos | sin]."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    ang = t.float()[:, None] * freqs[None]
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1)


# --------------------------------------------------------------------------------------
# UNet2DConditionModel (SD-1.4 unet/config.json)
# --------------------------------------------------------------------------------------
class _DownBlock(nn.Module):
    def __init__(self, cin, cout, temb, heads, cross_dim, with_attn, with_down):
        super().__init__()
        self.has_cross_attention = with_attn
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, temb, 1e-5) for i in range(2)])
        if with_attn:
            self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, cross_dim) for _ in range(2)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout, padding=1)]) if with_down else None

    def forward(self, x, temb, ctx):
        outs = []
        for i, res in enumerate(self.resnets):
            x = res(x, temb)
            if self.has_cross_attention:
                x = self.attentions[i](x, ctx)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class _MidBlock(nn.Module):
    def __init__(self, ch, temb, heads, cross_dim):
        super().__init__()
        self.attentions = nn.ModuleList([Transformer2DModel(ch, heads, cross_dim)])
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, temb, 1e-5) for _ in range(2)])

    def forward(self, x, temb, ctx):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, ctx)
        return self.resnets[1](x, temb)


class _UpBlock(nn.Module):
    def __init__(self, in_ch, prev_ch, out_ch, temb, heads, cross_dim, with_attn, with_up):
        super().__init__()
        self.has_cross_attention = with_attn
        res = []
        for j in range(3):
            skip = in_ch if j == 2 else out_ch
            rin = prev_ch if j == 0 else out_ch
            res.append(ResnetBlock2D(rin + skip, out_ch, temb, 1e-5))
        self.resnets = nn.ModuleList(res)
        if with_attn:
            self.attentions = nn.ModuleList([Transformer2DModel(out_ch, heads, cross_dim) for _ in range(3)])
        self.upsamplers = nn.ModuleList([Upsample2D(out_ch)]) if with_up else None


class UNet2DConditionModel(nn.Module):
    block_out_channels = (320, 640, 1280, 1280)

    def __init__(self, in_channels: int = 4, out_channels: int = 4, cross_dim: int = 768, heads: int = 8):
        super().__init__()
        ch = self.block_out_channels
        temb = ch[0] * 4
        self.conv_in = nn.Conv2d(in_channels, ch[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(ch[0], temb)
        self.down_blocks = nn.ModuleList()
        cout = ch[0]
        for i in range(4):
            cin, cout = cout, ch[i]
            self.down_blocks.append(_DownBlock(cin, cout, temb, heads, cross_dim, with_attn=i < 3, with_down=i < 3))
        self.mid_block = _MidBlock(ch[-1], temb, heads, cross_dim)
        rev = ch[::-1]
        self.up_blocks = nn.ModuleList()
        cout = rev[0]
        for i in range(4):
            prev, cout = cout, rev[i]
            cin = rev[min(i + 1, 3)]
            self.up_blocks.append(_UpBlock(cin, prev, cout, temb, heads, cross_dim, with_attn=i > 0, with_up=i < 3))
        self.conv_norm_out = nn.GroupNorm(32, ch[0], eps=1e-5)
        self.conv_out = nn.Conv2d(ch[0], out_channels, 3, padding=1)


def diffusion_unet(unet: UNet2DConditionModel, sample, timestep, encoder_hidden_states, res_time_embedding,
                   unet_block_indices: Sequence[int], unet_block_indices_type: str = "after", need_sample: bool = True):
    """Restates reference ``modeling/meta_arch/ldm_diffusers.py:454-616`` (+ :363-451 for up blocks)."""
    timesteps = timestep.expand(sample.shape[0])
    t_emb = timestep_sinusoid(timesteps, unet.block_out_channels[0])            # :498
    emb = unet.time_embedding(t_emb)                                            # :505
    if res_time_embedding is not None:                                          # :506-509
        if res_time_embedding.shape[1] == 1:
            res_time_embedding = res_time_embedding[:, 0]
        emb = emb + res_time_embedding
    sample = qr(unet.conv_in(q(sample)))                                         # :522
    skips = [sample]                                                            # :525
    for blk in unet.down_blocks:                                                # :526-538
        sample, outs = blk(sample, emb, encoder_hidden_states)
        skips += outs
    sample = unet.mid_block(sample, emb, encoder_hidden_states)                 # :552-559
    idx = 0
    feats = []
    for blk in unet.up_blocks:                                                  # :567-604
        for j, resnet in enumerate(blk.resnets):
            sample = torch.cat([sample, skips.pop()], dim=1)                    # :368-370 / :407-409
            if unet_block_indices_type == "in":
                if idx in unet_block_indices:
                    feats.append(sample)
                idx += 1
            sample = resnet(sample, emb)
            if blk.has_cross_attention:
                sample = blk.attentions[j](sample, encoder_hidden_states)
            if unet_block_indices_type == "after":                              # :389-392 / :442-445
                if idx in unet_block_indices:
                    feats.append(sample)
                idx += 1
        if blk.upsamplers is not None:
            sample = blk.upsamplers[0](sample)
    assert len(feats) == len(unet_block_indices)                                # :606
    out = None
    if need_sample:                                                             # :608-611
        out = unet.conv_out(q(F.silu(unet.conv_norm_out(sample))))
    return out, feats


# --------------------------------------------------------------------------------------
# AutoencoderKL encoder (SD-1.4 vae/config.json)
# --------------------------------------------------------------------------------------
class _EncDownBlock(nn.Module):
    def __init__(self, cin, cout, with_down):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, None, 1e-6) for i in range(2)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout, padding=0)]) if with_down else None


class _VaeMidBlock(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.attentions = nn.ModuleList([Attention(ch, heads=1, qkv_bias=True, group_norm=(32, 1e-6), residual=True)])
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, None, 1e-6) for _ in range(2)])

    def forward(self, x):
        x = self.resnets[0](x)
        x = self.attentions[0](x)
        return self.resnets[1](x)


class Encoder(nn.Module):
    def __init__(self, in_channels=3, latent=4, block_out=(128, 256, 512, 512)):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, block_out[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        cout = block_out[0]
        for i, c in enumerate(block_out):
            cin, cout = cout, c
            self.down_blocks.append(_EncDownBlock(cin, cout, with_down=i < len(block_out) - 1))
        self.mid_block = _VaeMidBlock(block_out[-1])
        self.conv_norm_out = nn.GroupNorm(32, block_out[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(block_out[-1], 2 * latent, 3, padding=1)


class _DecUpBlock(nn.Module):
    """diffusers UpDecoderBlock2D: 3 ResnetBlock2D (layers_per_block + 1, no time embedding) + nearest-2x Upsample2D."""

    def __init__(self, cin, cout, with_up):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, None, 1e-6) for i in range(3)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if with_up else None


class Decoder(nn.Module):
    """diffusers ``Decoder`` of SD-1.4 ``vae/config.json``: conv_in 4->512, mid block, four up blocks with output
    channels (512, 512, 256, 128), GN32(eps 1e-6) + SiLU + conv 128->3."""

    def __init__(self, out_channels=3, latent=4, block_out=(128, 256, 512, 512)):
        super().__init__()
        rev = block_out[::-1]
        self.conv_in = nn.Conv2d(latent, rev[0], 3, padding=1)
        self.mid_block = _VaeMidBlock(rev[0])
        self.up_blocks = nn.ModuleList()
        cout = rev[0]
        for i, c in enumerate(rev):
            cin, cout = cout, c
            self.up_blocks.append(_DecUpBlock(cin, cout, with_up=i < len(rev) - 1))
        self.conv_norm_out = nn.GroupNorm(32, block_out[0], eps=1e-6)
        self.conv_out = nn.Conv2d(block_out[0], out_channels, 3, padding=1)


class AutoencoderKL(nn.Module):
    """Encoder half (+quant_conv / post_quant_conv parameters) by default; ``with_decoder=True`` adds the decoder half
    used by the ``vae_decoder_loss`` / ``s0`` variant.  The decoder is constructed last so the
    random-init stream of the base configuration does not change."""
    scaling_factor = 0.18215
    latent_channels = 4

    def __init__(self, with_decoder: bool = False):
        super().__init__()
        self.encoder = Encoder()
        self.quant_conv = nn.Conv2d(8, 8, 1)
        self.post_quant_conv = nn.Conv2d(4, 4, 1)
        if with_decoder:
            self.decoder = Decoder()


@torch.no_grad()
def vae_encoder(vae: AutoencoderKL, images, encoder_block_indices: Sequence[int]):
    """Restates reference ``modeling/meta_arch/ldm_diffusers.py:283-311``."""
    index = 0
    features = []
    x = qr(vae.encoder.conv_in(q(images)))                                       # :287
    for blk in vae.encoder.down_blocks:                                         # :288-296
        for resnet in blk.resnets:
            x = resnet(x)
            index += 1
            if index in encoder_block_indices:
                features.append(x)
        if blk.downsamplers is not None:
            x = blk.downsamplers[0](x)
    x = vae.encoder.mid_block(x)                                                # :297
    x = q(F.silu(vae.encoder.conv_norm_out(x)))                                 # :299-300
    x = vae.encoder.conv_out(x)                                                 # :301
    moments = vae.quant_conv(x)                                                 # :303
    mean = moments[:, : vae.latent_channels]                                    # DiagonalGaussianDistribution.mean
    latents = mean * vae.scaling_factor                                         # :308
    assert len(encoder_block_indices) == len(features)
    return latents, features


@torch.no_grad()
def vae_decoder(vae: AutoencoderKL, latents, decoder_block_indices: Sequence[int], output_final: bool = False):
    """Restates reference ``modeling/meta_arch/ldm_diffusers.py:314-346``."""
    index = 0
    features = []
    latents = 1.0 / vae.scaling_factor * latents                                # :319
    sample = vae.post_quant_conv(latents)                                       # :320
    sample = qr(vae.decoder.conv_in(q(sample)))                                 # :322
    sample = vae.decoder.mid_block(sample)                                      # :325
    for blk in vae.decoder.up_blocks:                                           # :328-336
        for resnet in blk.resnets:
            if index in decoder_block_indices:
                features.append(sample)
            index += 1
            sample = resnet(sample)
        if blk.upsamplers is not None:
            sample = blk.upsamplers[0](sample)
    if output_final:                                                            # :339-342
        sample = vae.decoder.conv_out(q(F.silu(vae.decoder.conv_norm_out(sample))))
    else:
        sample = None
    return sample, features


# --------------------------------------------------------------------------------------
# DDPMScheduler.add_noise (SD-1.4 scheduler_config.json: scaled_linear 0.00085..0.012, 1000 steps)
# --------------------------------------------------------------------------------------
def ddpm_alphas_cumprod(num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012):
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def add_noise(latents, timesteps, shared_noise, alphas_cumprod=None):
    """Restates reference ``modeling/meta_arch/ldm_diffusers.py:349-360`` + DDPMScheduler.add_noise."""
    if alphas_cumprod is None:
        alphas_cumprod = ddpm_alphas_cumprod()
    if shared_noise.shape[2:] != latents.shape[2:]:                             # :351-353
        shared_noise = F.interpolate(shared_noise, size=latents.shape[2:], mode="bicubic", align_corners=False)
    noise = shared_noise.expand_as(latents)                                     # :356
    ac = alphas_cumprod.to(latents.device)
    sa = ac[timesteps].sqrt().view(-1, 1, 1, 1)
    sb = (1.0 - ac[timesteps]).sqrt().view(-1, 1, 1, 1)
    return sa * latents + sb * noise

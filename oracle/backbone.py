"""fp32 restatement of the MADM backbone boundary on top of oracle.sd14 / oracle.lora.

TEST INFRASTRUCTURE; PARITY UNPINNED (see oracle/__init__.py).  Follows:

* ``modeling/backbone/feature_extractor.py:20-284``   FeatureExtractorBackbone (single / slide forward)
* ``modeling/backbone/feature_extractor.py:287-396``  AttentionFeatureExtractorBackbone (+ GN bottleneck projections)
* ``modeling/meta_arch/ldm_base.py:632-717``          ClipFeatureProject
* ``modeling/meta_arch/ldm_base.py:720-924``          BasePromptTimeGenerator
* ``modeling/meta_arch/ldm_diffusers.py:17-217``      LdmDiffusers (constructor contract + forward)
* detectron2 ``BottleneckBlock(norm="GN")`` via ``ResNet.make_stage``

Device-agnostic (the reference hard-codes ``.cuda()``); no CLIP text tower: ``uncond_inputs`` is a
seeded stand-in for CLIP('') (reference ``ldm_diffusers.py:76,219-243``), as SURVEY §8d specifies.
"""
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import sd14
from .sd14 import q


# ----------------------------------------------------------------------------- detectron2 pieces
class D2Conv2d(nn.Conv2d):
    """detectron2.layers.Conv2d: conv -> norm -> activation, with ``.norm`` as a child module."""

    def __init__(self, cin, cout, k, padding=0, norm: Optional[nn.Module] = None):
        super().__init__(cin, cout, k, padding=padding, bias=False)
        self.norm = norm

    def forward(self, x):
        x = F.conv2d(x, self.weight, None, self.stride, self.padding)
        return self.norm(x) if self.norm is not None else x


class BottleneckBlock(nn.Module):
    """detectron2 BottleneckBlock(in, out, bottleneck_channels, stride=1, norm='GN'); GN32 eps 1e-5;
    c2_msra_fill init (kaiming_normal fan_out relu)."""

    def __init__(self, cin: int, cout: int, bottleneck_channels: int):
        super().__init__()
        gn = lambda c: nn.GroupNorm(32, c)  # noqa: E731
        self.shortcut = D2Conv2d(cin, cout, 1, norm=gn(cout)) if cin != cout else None
        self.conv1 = D2Conv2d(cin, bottleneck_channels, 1, norm=gn(bottleneck_channels))
        self.conv2 = D2Conv2d(bottleneck_channels, bottleneck_channels, 3, padding=1, norm=gn(bottleneck_channels))
        self.conv3 = D2Conv2d(bottleneck_channels, cout, 1, norm=gn(cout))
        for layer in (self.conv1, self.conv2, self.conv3, self.shortcut):
            if layer is not None:
                nn.init.kaiming_normal_(layer.weight, mode="fan_out", nonlinearity="relu")

    def forward(self, x):
        x = q(x)
        out = q(F.relu(self.conv1(x)))
        out = q(F.relu(self.conv2(out)))
        out = self.conv3(out)
        sc = self.shortcut(x) if self.shortcut is not None else x
        return F.relu(out + sc)


def make_projection(cin: int, cout: int, bottleneck_channels: int, num_blocks: int = 1) -> nn.Sequential:
    blocks = []
    for i in range(num_blocks):
        blocks.append(BottleneckBlock(cin if i == 0 else cout, cout, bottleneck_channels))
    return nn.Sequential(*blocks)


# ----------------------------------------------------------------------------- conditioning
def trunc_normal_(t, std):
    return nn.init.trunc_normal_(t, std=std, a=-2.0, b=2.0)  # timm trunc_normal_ defaults


class ClipFeatureProject(nn.Module):
    """ldm_base.py:632-717 for ``input_prefix=False`` (clip_state='no', the shipped config)."""

    def __init__(self, learnable_cond_prompt, prompt_out_features, prompt_seq_len, learnable_cond_time,
                 time_out_features, time_seq_len, time_alpha_cond_size, without_prompt_alpha=False):
        super().__init__()
        self.learnable_cond_prompt = learnable_cond_prompt
        self.learnable_cond_time = learnable_cond_time
        self.without_prompt_alpha = without_prompt_alpha
        if learnable_cond_prompt:
            pe = torch.zeros(1, prompt_seq_len, prompt_out_features)
            trunc_normal_(pe, std=0.02)                                              # :653
            self.prompt_embed = nn.Parameter(pe)
            if not without_prompt_alpha:
                shape = [1, prompt_seq_len, prompt_out_features]
                self.alpha_cond_prompt = nn.Parameter(torch.rand(shape))             # :664
                self.alpha_uncond_prompt = nn.Parameter(torch.rand(shape))           # :665
        if learnable_cond_time:
            self.alpha_cond_time = nn.Parameter(torch.zeros(time_alpha_cond_size))   # :668
            te = torch.zeros(1, time_seq_len, time_out_features)
            trunc_normal_(te, std=0.02)                                              # :670-671
            self.time_embed = nn.Parameter(te)

    def forward(self, uncond_prompt, prefix=None):
        if self.learnable_cond_prompt:
            if not self.without_prompt_alpha:                                        # :681
                cond_prompt = torch.tanh(self.alpha_uncond_prompt) * uncond_prompt + \
                    torch.tanh(self.alpha_cond_prompt) * self.prompt_embed
            else:
                cond_prompt = self.prompt_embed
        else:
            cond_prompt = uncond_prompt
        cond_time = torch.tanh(self.alpha_cond_time) * self.time_embed if self.learnable_cond_time else None  # :706
        return cond_prompt, cond_time


class LdmDiffusers(nn.Module):
    """ldm_diffusers.py:17-217 with random-init SD-1.4 modules instead of ``from_pretrained``."""
    latent_image_size = (64, 64)
    text_embed_shape = torch.Size([77, 768])
    unet_time_embed_out_features = 1280
    feature_dims = [512, 512, 2560, 1920, 960, 640, 512, 512]
    feature_strides = [4, 8, 64, 32, 16, 8, 8, 4]
    num_groups = 8
    grouped_indices = [[0], [1], [2], [3], [4], [5], [6], [7]]
    input_mean = 0.5
    input_std = 0.5

    def __init__(self, stable_diffusion_name_or_path=None, encoder_block_indices=(5,), unet_block_indices=(5, 8, 11),
                 decoder_block_indices=(), input_range="01", unet_block_indices_type="in", finetune_unet="no",
                 vae_decoder_loss=False, final_fuse_vae_decoder_feat=False, **unused):
        super().__init__()
        assert input_range in {"01", "-1+1"}
        assert unet_block_indices_type in {"in", "after"}
        self.encoder_block_indices = list(encoder_block_indices)
        self.unet_block_indices = list(unet_block_indices)
        self.decoder_block_indices = list(decoder_block_indices)
        self.input_range = input_range
        self.unet_block_indices_type = unet_block_indices_type
        self.finetune_unet = finetune_unet
        self.vae_decoder_loss = vae_decoder_loss
        self.final_fuse_vae_decoder_feat = final_fuse_vae_decoder_feat
        self.vae = sd14.AutoencoderKL()
        self.unet = sd14.UNet2DConditionModel()
        if vae_decoder_loss or self.decoder_block_indices:  # built after the base modules: the base random-init stream is unchanged
            self.vae.decoder = sd14.Decoder()
        self.register_buffer("alphas_cumprod", sd14.ddpm_alphas_cumprod(), persistent=False)
        rng = torch.Generator().manual_seed(42)                                       # :73-75
        self.register_buffer("shared_noise", torch.randn(1, 4, *self.latent_image_size, generator=rng))
        rng7 = torch.Generator().manual_seed(7)                                       # stand-in for CLIP('') :76
        self.register_buffer("uncond_inputs", torch.randn(1, 77, 768, generator=rng7))

    def forward(self, batched_inputs, input_modal, **kwargs):
        images = batched_inputs["img"]
        if self.input_range == "-1+1":                                                # :145-147
            images = (images - self.input_mean) / self.input_std
            assert -1 <= torch.min(images) and torch.max(images) <= 1
        text_prompt = batched_inputs["cond_inputs"]
        res_time_embedding = batched_inputs["cond_emb"]
        latents, encoder_features = sd14.vae_encoder(self.vae, images, self.encoder_block_indices)   # :151
        bsz = latents.shape[0]
        lo, hi = batched_inputs["timestep"] if "timestep" in batched_inputs else (0, 1)              # :156-159
        if "timesteps_override" in batched_inputs:      # test hook: fixed per-sample timesteps
            timesteps = batched_inputs["timesteps_override"].long()
        else:
            timesteps = torch.randint(low=lo, high=hi, size=(bsz,), device=latents.device).long()    # :160
        noisy = sd14.add_noise(latents, timesteps, self.shared_noise, self.alphas_cumprod)           # :162
        forward_unet = self.unet
        if kwargs.get("ema_forward") and hasattr(self, "ema_unet"):                                  # :182-185
            forward_unet = self.ema_unet
        sample, unet_features = sd14.diffusion_unet(forward_unet, noisy, timesteps, text_prompt, res_time_embedding,
                                                    self.unet_block_indices, self.unet_block_indices_type,
                                                    need_sample=self.vae_decoder_loss)               # :186
        self.last_intermediates = dict(latents=latents, noisy_latents=noisy, timesteps=timesteps)
        decoder_features = []
        if self.vae_decoder_loss:                                                                    # :191-201
            decoder_output, _ = sd14.vae_decoder(self.vae, sample, [], output_final=True)
            if self.final_fuse_vae_decoder_feat:
                decoder_features = [decoder_output.detach()]
            else:
                assert len(self.encoder_block_indices) == 0
                encoder_features = [decoder_output.detach()]
        elif self.decoder_block_indices:                                                             # :203-205
            _, decoder_features = sd14.vae_decoder(self.vae, latents, self.decoder_block_indices)
        feats = [*encoder_features, *unet_features, *decoder_features]
        if kwargs.get("return_unet_final_output"):                                                   # :211-215
            return feats, {"before_vae.decoder": sample, "after_vae.decoder": torch.clip(decoder_output, min=-1.0, max=1.0)}
        return feats                                                                                 # :217


class BasePromptTimeGenerator(nn.Module):
    """ldm_base.py:720-924 (clip_state='no')."""

    def __init__(self, learnable_cond_prompt=True, learnable_cond_time=True, same_cond_params=False,
                 detach_prompt_for_mixed_data=False, clip_state="no", num_timesteps=1, clip_model_name="",
                 ldm_extractor=None, without_prompt_alpha=False, multi_layer_prompt=False,
                 mix_source_target_prompt=False, init_uncond_prompt=False, mask_prompt_ratio=False,
                 detach_mask_prompt=False, prompt_perturbation=False, rand_prompt_scale=None, **kwargs):
        super().__init__()
        assert clip_state == "no" and not multi_layer_prompt and not init_uncond_prompt
        self.same_cond_params = same_cond_params
        self.mix_source_target_prompt = mix_source_target_prompt
        self.mask_prompt_ratio = mask_prompt_ratio
        self.prompt_perturbation = prompt_perturbation
        self.rand_prompt_scale = rand_prompt_scale
        self.ldm_extractor = ldm_extractor
        self.text_embed_shape = ldm_extractor.text_embed_shape
        t_out = ldm_extractor.unet_time_embed_out_features
        mk = lambda: ClipFeatureProject(learnable_cond_prompt, self.text_embed_shape[1], self.text_embed_shape[0],  # noqa: E731
                                        learnable_cond_time, t_out, num_timesteps, t_out, without_prompt_alpha)
        self.clip_project_rgb = mk()                                                  # :790-806
        self.clip_project_others = self.clip_project_rgb if same_cond_params else mk()  # :811-830

    @property
    def uncond_inputs(self):
        return self.ldm_extractor.uncond_inputs

    def forward(self, batched_inputs, input_modal, ema_forward=False, timestep=None, **kwargs):
        assert input_modal in {"rgb", "others", "mixed", "masked_prompt", "prompt_perturbation", "rand_prompt"}
        image = batched_inputs["img"]
        if input_modal == "rgb":                                                      # :877-879
            assert ema_forward is False
            ci, ce = self.clip_project_rgb(self.uncond_inputs)
        elif input_modal == "mixed" and self.mix_source_target_prompt:                # :880-884
            s_ci, s_ce = self.clip_project_rgb(self.uncond_inputs)
            t_ci, t_ce = self.clip_project_others(self.uncond_inputs)
            ci, ce = (s_ci + t_ci) / 2, (s_ce + t_ce) / 2
        else:                                                                         # :885-887
            proj = self.ema_clip_project_others if ema_forward else self.clip_project_others
            ci, ce = proj(self.uncond_inputs)
        if input_modal == "masked_prompt" and self.mask_prompt_ratio:                 # :892-897
            mask = (torch.rand((1, ci.shape[0], ci.shape[1], 1), device=ci.device) > self.mask_prompt_ratio).float()
            ci = ci * mask[0]
        elif input_modal == "prompt_perturbation" and self.prompt_perturbation:       # :898-901
            ci = (ci + torch.randn(ci.shape, device=ci.device) * self.prompt_perturbation).detach()
        elif input_modal == "rand_prompt":                                            # :902-903
            ci = torch.rand_like(ci) * self.rand_prompt_scale
        batched_inputs["cond_inputs"], batched_inputs["cond_emb"] = ci, ce
        if timestep is not None:                                                      # :911-912
            batched_inputs["timestep"] = timestep
        if image.shape[0] != 1:                                                       # :915-917
            batched_inputs["cond_inputs"] = torch.repeat_interleave(ci, repeats=image.shape[0], dim=0)
            batched_inputs["cond_emb"] = torch.repeat_interleave(ce, repeats=image.shape[0], dim=0)
        return self.ldm_extractor(batched_inputs, input_modal, ema_forward=ema_forward, **kwargs)


# ----------------------------------------------------------------------------- backbone
class AttentionFeatureExtractorBackbone(nn.Module):
    """feature_extractor.py:287-396 on top of FeatureExtractorBackbone (:20-284)."""

    def __init__(self, attention_features_res=None, feature_dims=(512, 320, 640, 1280), attention_features_location=None,
                 target_attention_loss=False, attention_select_index=None, feature_extractor=None,
                 out_features=("s2", "s3", "s4", "s5"), backbone_in_size=(512, 512), min_stride=4, max_stride=32,
                 projection_dim=(512, 512, 512, 512), bottleneck_channels=128, num_res_blocks=1, use_checkpoint=False,
                 slide_training=False, slide_inference=False):
        super().__init__()
        self.feature_extractor = feature_extractor
        self.feature_dims = list(feature_dims)
        self.backbone_in_size = tuple(backbone_in_size)
        self._slide_inference = slide_inference
        self._slide_training = slide_training
        self.target_attention_loss = target_attention_loss
        self.feature_projections = nn.ModuleList(
            [make_projection(fd, projection_dim[i], bottleneck_channels, num_res_blocks)
             for i, fd in enumerate(self.feature_dims)])                              # :347-359
        self._out_feature_strides = {s: 2 ** int(s[1]) for s in out_features}         # :361-364
        self._out_feature_channels = {s: projection_dim[i] for i, s in enumerate(out_features)}
        self._out_features = list(out_features)

    size_divisibility = 64                                                            # :127-129

    def preprocess_image(self, img):                                                  # :140-146
        if not self._slide_inference and tuple(img.shape[-2:]) != self.backbone_in_size:
            # T.Resize(..., BILINEAR) on a tensor: torchvision 0.16.1 (README.md:31) defaults antialias to "warn" = no antialiasing
            img = F.interpolate(img, size=self.backbone_in_size, mode="bilinear", align_corners=False, antialias=False)
        h, w = img.shape[-2:]
        ph, pw = (-h) % 64, (-w) % 64
        return F.pad(img, (0, pw, 0, ph)) if (ph or pw) else img

    def forward_features(self, features, input_image_size=None, ema_forward=False):   # :367-396
        by_width = {f.shape[-1]: f for f in features}                                 # :371-373
        out = {}
        for idx, name in enumerate(self._out_features):
            res = 512 // self._out_feature_strides[name]                              # :382-385
            proj = self.ema_feature_projections[idx] if ema_forward else self.feature_projections[idx]
            out[name] = proj(by_width[res])
        return {"output_features": out}

    def single_forward(self, img, input_modal="rgb", ema_forward=False, timestep=None, **kwargs):  # :156-170
        size = img.shape[-2:]
        img = self.preprocess_image(img)
        feats = self.feature_extractor(dict(img=img), input_modal, ema_forward, timestep, **kwargs)
        if "return_unet_final_output" in kwargs:                                       # :164-166
            return self.forward_features(feats[0], size, ema_forward), feats[1]
        return self.forward_features(feats, size, ema_forward)

    def slide_windows(self, h_img, w_img, crop=512, stride=256):
        """512^2 crops at stride 256 on both axes; on 512x1024 this is exactly the
        reference's three hard-coded windows (feature_extractor.py:75)."""
        ys = list(range(0, max(h_img - crop, 0) + 1, stride))
        xs = list(range(0, max(w_img - crop, 0) + 1, stride))
        return [(y, y + crop, x, x + crop) for y in ys for x in xs]

    def slide_forward(self, img, input_modal="rgb", ema_forward=False, timestep=None, **kwargs):   # :199-278
        b, _, h_img, w_img = img.shape
        outs = {k: torch.zeros((b, self._out_feature_channels[k], h_img // s, w_img // s), dtype=img.dtype, device=img.device)
                for k, s in self._out_feature_strides.items()}
        cnt = {k: torch.zeros_like(v) for k, v in outs.items()}
        for (y1, y2, x1, x2) in self.slide_windows(h_img, w_img):
            crop = self.single_forward(img[:, :, y1:y2, x1:x2], input_modal, ema_forward, timestep, **kwargs)["output_features"]
            for k, s in self._out_feature_strides.items():                            # :255-271
                outs[k][:, :, y1 // s:y2 // s, x1 // s:x2 // s] += crop[k]
                cnt[k][..., y1 // s:y2 // s, x1 // s:x2 // s] += 1
        for k in outs:                                                                # :274-275
            outs[k] /= cnt[k]
        return {"output_features": outs}

    def forward(self, img, input_modal="rgb", ema_forward=False, timestep=None, **kwargs):         # :280-284
        if (self.training and not self._slide_training) or not self._slide_inference:
            return self.single_forward(img, input_modal, ema_forward, timestep, **kwargs)
        return self.slide_forward(img, input_modal, ema_forward, timestep, **kwargs)

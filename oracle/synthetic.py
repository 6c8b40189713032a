"""Seeded synthetic model + inputs shared by the parity tests, the golden-fixture generator and
the CPU baseline (SURVEY §8d / BASELINE.md §3).  TEST INFRASTRUCTURE.

Weights: PyTorch default inits under ``torch.manual_seed(1234)``; the zero-initialised pieces of
the reference are de-degenerated so they are exercised: ``lora_B ~ N(0, 0.02)`` (reference init 0,
``mtmadise.py:122``), ``alpha_cond_time ~ U(0,1)`` (reference init 0, ``ldm_base.py:668``).  On top of
SURVEY's recipe every norm affine is perturbed (weight 1+0.1N, bias 0.1N) so gamma/beta handling is
tested too.
"""
import copy
from typing import Sequence

import torch
import torch.nn as nn

from . import backbone as ob
from .lora import LoraLinear, add_adapter, set_adapter

CONFIG = dict(  # reference config_files/common/models/mtmadise_multi_lora.py:14-41
    feature_dims=[512, 320, 640, 1280], projection_dim=[512, 512, 512, 512],
    encoder_block_indices=[5], unet_block_indices=[5, 8, 11], unet_block_indices_type="after",
    decoder_block_indices=(), input_range="-1+1", out_features=["s2", "s3", "s4", "s5"],
)
# the variant all three shipped experiment files select (config_files/SemSeg/MTMADISE/mtmadise_cityscapes_rgb_to_depth_11.py:47-55):
# UNet final output -> VAE decoder -> 3-channel 512^2 "feature" -> Bottleneck(3 -> 128 -> 128) -> 's0'   
CONFIG_S0 = dict(CONFIG, feature_dims=[3, 320, 640, 1280], projection_dim=[128, 512, 512, 512], encoder_block_indices=[],
                 out_features=["s0", "s3", "s4", "s5"], vae_decoder_loss=True)


def build_backbone(lora_configs: Sequence[str] = ("default_r16_a16", "Depth_r16_a16"), seed: int = 1234,
                   same_cond_params: bool = False, with_ema: bool = True, variant: str = "base") -> ob.AttentionFeatureExtractorBackbone:
    CONFIG = globals()["CONFIG"] if variant == "base" else CONFIG_S0
    torch.manual_seed(seed)
    ldm = ob.LdmDiffusers(
        stable_diffusion_name_or_path=None, encoder_block_indices=CONFIG["encoder_block_indices"],
        unet_block_indices=CONFIG["unet_block_indices"], unet_block_indices_type=CONFIG["unet_block_indices_type"],
        decoder_block_indices=CONFIG["decoder_block_indices"], input_range=CONFIG["input_range"], finetune_unet="all",
        vae_decoder_loss=CONFIG.get("vae_decoder_loss", False))
    gen = ob.BasePromptTimeGenerator(learnable_cond_prompt=True, learnable_cond_time=True, clip_state="no",
                                     num_timesteps=1, clip_model_name="ViT-L-14-336", ldm_extractor=ldm,
                                     same_cond_params=same_cond_params)
    bb = ob.AttentionFeatureExtractorBackbone(
        attention_features_res=None, feature_dims=CONFIG["feature_dims"], projection_dim=CONFIG["projection_dim"],
        attention_features_location=None, feature_extractor=gen, num_res_blocks=1,
        out_features=CONFIG["out_features"], use_checkpoint=False, slide_training=False)
    for cfg in lora_configs:  # MTMADISE.__init__ parsing, mtmadise.py:48-54
        name, rank, alpha = cfg.split("_")
        add_adapter(ldm.unet, name, int(rank[1:]), int(alpha[1:]))
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for m in bb.modules():
            if isinstance(m, LoraLinear):
                for a in m.lora_B:
                    m.lora_B[a].weight.copy_(torch.randn(m.lora_B[a].weight.shape, generator=g) * 0.02)
            if isinstance(m, ob.ClipFeatureProject) and m.learnable_cond_time:
                m.alpha_cond_time.copy_(torch.rand(m.alpha_cond_time.shape, generator=g))
            if isinstance(m, (nn.GroupNorm, nn.LayerNorm)):
                m.weight.copy_(1.0 + 0.1 * torch.randn(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
    if with_ema:  # CMDISE._inti_ema_weights, cmdise.py:307-325 (deep copies), perturbed so EMA != student
        bb.ema_feature_projections = copy.deepcopy(bb.feature_projections)
        gen.ema_clip_project_others = copy.deepcopy(gen.clip_project_others)
        with torch.no_grad():
            for p in list(bb.ema_feature_projections.parameters()) + list(gen.ema_clip_project_others.parameters()):
                p.add_(0.01 * torch.randn(p.shape, generator=g))
    if lora_configs:
        set_adapter(ldm.unet, [lora_configs[-1].split("_")[0]])
    bb.eval()
    for p in bb.parameters():
        p.requires_grad_(False)
    return bb


def training_gradients(bb, img: torch.Tensor, adapter: str = "Depth", input_modal: str = "others", seed: int = 99):
    """Reference gradients for SURVEY §8 row f-3 / BASELINE config 5 (the target of the backward pass that is still to be built): the
    trainable set of the LoRA training step — the active adapter's A / B factors, the feature projections and the learned prompt / time
    parameters — receives the gradient of a fixed linear functional of the feature dict, L = sum_k <feat_k, R_k> / 1000 with seeded
    Gaussian R_k.  As in the reference the VAE encoder runs under no_grad (ldm_diffusers.py:282), so gradients flow through the UNet and
    the projections only.  Returns (loss, {name: grad}) with None for parameters the path does not reach."""
    from .lora import set_adapter
    set_adapter(bb.feature_extractor.ldm_extractor.unet, [adapter])
    train = [(n, p) for n, p in bb.named_parameters()
             if ("lora_" in n and f".{adapter}." in n) or n.startswith("feature_projections.") or "clip_project_others" in n]
    for p in bb.parameters():
        p.requires_grad_(False)
        p.grad = None
    for _, p in train:
        p.requires_grad_(True)
    out = bb(img, input_modal=input_modal)["output_features"]
    g = torch.Generator().manual_seed(seed)
    loss = sum((v * torch.randn(v.shape, generator=g).to(v.device)).sum() for v in out.values()) / 1e3
    loss.backward()
    grads = {n: (None if p.grad is None else p.grad.detach().clone()) for n, p in train}
    for _, p in train:
        p.requires_grad_(False)
        p.grad = None
    return float(loss.detach()), grads


def synthetic_images(batch: int, h: int = 512, w: int = 512, seed: int = 0) -> torch.Tensor:
    """``img255 = rand*255`` as the dataloader would give; the meta-arch divides by pixel_std=255."""
    g = torch.Generator().manual_seed(seed)
    img255 = torch.rand(batch, 3, h, w, generator=g) * 255.0
    return img255 / 255.0

"""Drop-in mirrors of the reference's ``LdmDiffusers`` (``modeling/meta_arch/ldm_diffusers.py:17-217``) and
``BasePromptTimeGenerator`` / ``ClipFeatureProject`` (``modeling/meta_arch/ldm_base.py:632-924``).

Same constructor kwargs, attribute names and state_dict keys; the forward pass is the CUDA engine
(``madm_extract``), not diffusers.  The tiny batch-invariant conditioning arithmetic
(``tanh(alpha) * embed`` on ``[1,77,768]`` / ``[1,1,1280]``) stays in PyTorch, as SURVEY §8 a-2 prescribes.
"""
import logging
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib
from .engine import Engine
from .sd14_params import UNetParams, VAEParams

logger = logging.getLogger(__name__)


def _load_sd_component(root: str, sub: str) -> Optional[Dict[str, torch.Tensor]]:
    """Read ``<root>/<sub>/diffusion_pytorch_model.{safetensors,bin}`` of a local SD-1.4 snapshot, if present."""
    d = os.path.join(root, sub)
    st = os.path.join(d, "diffusion_pytorch_model.safetensors")
    if os.path.exists(st):
        from safetensors.torch import load_file
        return load_file(st)
    pt = os.path.join(d, "diffusion_pytorch_model.bin")
    if os.path.exists(pt):
        return torch.load(pt, map_location="cpu")
    return None


class FeatureTaps(list):
    """What ``LdmDiffusers.forward`` returns: the reference's ``[*encoder_features, *unet_features]`` list
    (``ldm_diffusers.py:217``) — NCHW fp32 tensors — plus the engine token that lets the backbone run the
    projection stage on the workspace-resident NHWC taps without a round trip."""
    token: Optional[Tuple] = None


class LdmDiffusers(nn.Module):
    latent_image_size = (64, 64)
    text_embed_shape = torch.Size([77, 768])
    unet_time_embed_out_features = 1280
    uncond_inputs_size = torch.Size([1, 77, 768])
    feature_size = (512, 512)
    feature_dims = [512, 512, 2560, 1920, 960, 640, 512, 512]
    feature_strides = [4, 8, 64, 32, 16, 8, 8, 4]
    num_groups = 8
    grouped_indices = [[0], [1], [2], [3], [4], [5], [6], [7]]
    timesteps = 0
    input_mean = 0.5
    input_std = 0.5

    def __init__(self, stable_diffusion_name_or_path, encoder_block_indices, unet_block_indices, decoder_block_indices,
                 input_range="01", unet_block_indices_type="in", finetune_unet="no", concat_pixel_shuffle=False,
                 add_latent_noise=-1, norm_latent_noise=False, vae_decoder_loss=False, input_channel_plus=0,
                 final_fuse_vae_decoder_feat=False, device=None, uncond_inputs: Optional[torch.Tensor] = None,
                 compute_dtype: str = "fp16"):
        super().__init__()
        self.stable_diffusion_name_or_path = os.path.expanduser(stable_diffusion_name_or_path) if stable_diffusion_name_or_path else None
        self.encoder_block_indices = list(encoder_block_indices)
        self.unet_block_indices = list(unet_block_indices)
        self.decoder_block_indices = list(decoder_block_indices)
        self.input_range = input_range
        assert self.input_range in {"01", "-1+1"}
        self.unet_block_indices_type = unet_block_indices_type
        assert self.unet_block_indices_type in {"in", "after"}
        self.finetune_unet = finetune_unet
        assert self.finetune_unet in {"no", "all", "attention", "without cross-attention"}
        self.add_latent_noise = add_latent_noise
        self.norm_latent_noise = norm_latent_noise
        self.vae_decoder_loss = vae_decoder_loss
        self.final_fuse_vae_decoder_feat = final_fuse_vae_decoder_feat
        # rows of SURVEY §8 that are "next" / out of scope fail loudly instead of silently computing something else
        unsupported = []
        # two supported configurations: the base one (encoder tap -> s2) and the shipped experiments' s0 variant
        # (vae_decoder_loss=True, encoder_block_indices=[]: decoded image -> s0)
        self.variant = "s0" if vae_decoder_loss else "base"
        if self.encoder_block_indices != ([] if vae_decoder_loss else [5]):
            unsupported.append("encoder_block_indices must be [5] (base) or [] with vae_decoder_loss=True (ldm_diffusers.py:198)")
        if self.unet_block_indices != [5, 8, 11] or unet_block_indices_type != "after": unsupported.append("unet taps != [5,8,11]/'after'")
        if self.decoder_block_indices: unsupported.append("decoder_block_indices")
        if input_range != "-1+1": unsupported.append("input_range='01'")
        if concat_pixel_shuffle or input_channel_plus or norm_latent_noise or add_latent_noise != -1:
            unsupported.append("concat_pixel_shuffle / input_channel_plus / latent-noise variants")
        if final_fuse_vae_decoder_feat: unsupported.append("final_fuse_vae_decoder_feat")
        if unsupported:
            raise NotImplementedError("madm_b200 implements the base hot-path configuration "
                                      "(config_files/common/models/mtmadise_multi_lora.py:14-41) and its vae_decoder_loss / s0 variant "
                                      "(config_files/SemSeg/MTMADISE/mtmadise_cityscapes_rgb_to_depth_11.py:47-55); unsupported: "
                                      + ", ".join(unsupported))
        device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.vae = VAEParams(device=device, with_decoder=bool(vae_decoder_loss))
        self.unet = UNetParams(device=device)
        import weakref
        object.__setattr__(self.vae, "_owner", weakref.ref(self))  # lets the module-level vae_encoder(vae, ...) find the engine
        loaded = False
        if self.stable_diffusion_name_or_path and os.path.isdir(self.stable_diffusion_name_or_path):
            for sub, mod in (("vae", self.vae), ("unet", self.unet)):
                sd = _load_sd_component(self.stable_diffusion_name_or_path, sub)
                if sd is not None:
                    own = mod.state_dict()
                    mod.load_state_dict({k: v.float() for k, v in sd.items() if k in own}, strict=False)
                    loaded = True
        if not loaded:
            logger.warning("SD-1.4 snapshot not found at %s: UNet/VAE are randomly initialised (load a MADM checkpoint "
                           "with load_state_dict to set them)", self.stable_diffusion_name_or_path)
        rng = torch.Generator().manual_seed(42)  # reference ldm_diffusers.py:73-75 (CPU generator -> bit-reproducible)
        self.register_buffer("shared_noise", torch.randn(1, 4, *self.latent_image_size, generator=rng).to(device))
        if uncond_inputs is None:
            # CLIP('') embedding (ldm_diffusers.py:76,219-243) comes from the checkpoint buffer; until one is loaded a seeded
            # stand-in keeps the module usable (SURVEY §8d synthetic recipe)
            uncond_inputs = torch.randn(1, 77, 768, generator=torch.Generator().manual_seed(7))
        self.register_buffer("uncond_inputs", uncond_inputs.detach().to(device))
        self.compute_dtype = compute_dtype  # 'fp16' (reference AMP dtype, default) or 'bf16'; see DESIGN.md Numerics
        self._engine: Optional[Engine] = None
        self._named_cache: Dict = {}
        self._bind_cache = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_parameter_cache())
        self._ema_engine: Optional[Engine] = None  # second context for `ema_unet` (CMDISE ema_w_unet, cmdise.py:318-321)
        self._last_ema_unet = False
        self._bound_extra: List[Tuple[str, torch.Tensor]] = []
        self._serial = 0
        self._freeze()

    # ---------------------------------------------------------------------------- reference surface
    def _freeze(self):  # ldm_diffusers.py:101-121
        super().train(mode=False)
        for p in self.parameters():
            p.requires_grad = False
        if self.finetune_unet != "no":
            for name, p in self.unet.named_parameters():
                if self.finetune_unet == "all":
                    p.requires_grad = True
                elif self.finetune_unet == "attention" and "attentions" in name:
                    p.requires_grad = True
                elif self.finetune_unet == "without cross-attention" and not ("attentions" in name and "attn2" in name):
                    p.requires_grad = True

    def train(self, mode: bool = True):
        super().train(False)
        return self

    @property
    def device(self):
        return self.shared_noise.device

    # ---------------------------------------------------------------------------- engine plumbing
    def engine(self, ema_unet: bool = False) -> Engine:
        if ema_unet:
            if self._ema_engine is None:
                self._ema_engine = Engine(self.device, self.compute_dtype, self.variant)
            return self._ema_engine
        if self._engine is None:
            self._engine = Engine(self.device, self.compute_dtype, self.variant)
        return self._engine

    def named_engine_tensors(self, ema_unet: bool = False) -> List[Tuple[str, torch.Tensor]]:
        """(state_dict key, Parameter) of every UNet / VAE parameter.  Walking the ~1500-parameter module tree costs milliseconds, so the
        list is cached per (UNet object, its structure version); the Parameters themselves are stored, so moved storage (`.to()`,
        `load_state_dict`) is seen by the engine's data_ptr check.  `invalidate_parameter_cache()` drops it (done automatically after
        `load_state_dict`, which may replace Parameter objects with assign=True)."""
        unet = self.ema_unet if ema_unet else self.unet  # the EMA teacher's UNet is bound under the same names in its own context
        key = (bool(ema_unet), id(unet), getattr(unet, "_struct_version", 0))
        hit = self._named_cache.get(bool(ema_unet))
        if hit is not None and hit[0] == key:
            return hit[1]
        pre = "feature_extractor.ldm_extractor."
        out = [(pre + "unet." + n, p) for n, p in unet.named_parameters()]
        out += [(pre + "vae." + n, p) for n, p in self.vae.named_parameters()]
        self._named_cache[bool(ema_unet)] = (key, out)
        return out

    def invalidate_parameter_cache(self):
        self._named_cache.clear()
        self._bind_cache = None

    def prepare(self, extra: Sequence[Tuple[str, torch.Tensor]] = (), ema_unet: bool = False):
        """Bind parameter pointers and (re)pack weights if anything changed (version counters, adapter switch)."""
        eng = self.engine(ema_unet)
        if extra:  # remembered, so callers without the projection tensors (forward(), vae_encoder()) do not force a re-bind / repack
            self._bound_extra = extra if isinstance(extra, list) else list(extra)
        base = self.named_engine_tensors(ema_unet)
        bc = self._bind_cache
        if bc is None or bc[0] is not base or bc[1] is not self._bound_extra:  # the concatenation is cached with its parts
            bc = (base, self._bound_extra, base + list(self._bound_extra))
            self._bind_cache = bc
        eng.bind(bc[2])
        unet = self.ema_unet if ema_unet else self.unet
        adapter = unet.active_adapter()
        eng.ensure_packed(adapter, unet.scaling_of(adapter) if adapter else 0.0)
        return eng

    def sample_timesteps(self, batched_inputs, bsz: int) -> torch.Tensor:
        lo, hi = batched_inputs["timestep"] if "timestep" in batched_inputs else (0, 1)  # ldm_diffusers.py:156-161
        if not 0 <= int(lo) < int(hi) <= 1000:  # q-sample indexes the 1000-entry alpha-bar table of DDPMScheduler with these
            raise ValueError(f"timestep range must satisfy 0 <= lo < hi <= 1000 (num_train_timesteps), got {(lo, hi)}")
        return torch.randint(low=int(lo), high=int(hi), size=(bsz,), device=self.device).long()

    def run(self, batched_inputs, input_modal, *, stages, ema_projections=False, extra=(), want_taps=False, want_latents=False,
            timesteps: Optional[torch.Tensor] = None, out=None, out_dtype=torch.float32, **kwargs):
        use_ema_unet = bool(kwargs.get("ema_forward")) and hasattr(self, "ema_unet")  # ldm_diffusers.py:182-185
        if torch.is_grad_enabled() and any(p.requires_grad for p in (self.ema_unet if use_ema_unet else self.unet).parameters()):
            raise NotImplementedError(
                "LdmDiffusers called under torch.enable_grad() with trainable UNet parameters: the engine's taps carry no grad_fn. "
                "Call the backbone (AttentionFeatureExtractorBackbone.forward), whose training path back-propagates through the engine, "
                "or wrap inference in torch.no_grad()")
        if "modality_mask" in kwargs:
            raise NotImplementedError("modality_mask needs input_channel_plus != 0, which is outside the shipped configs")
        want_final = bool(kwargs.get("return_unet_final_output"))
        if want_final and self.variant != "s0":
            raise NotImplementedError("return_unet_final_output needs vae_decoder_loss=True (the reference raises on decoder_output too)")
        images = batched_inputs["img"]
        bsz = images.shape[0]
        eng = self.prepare(extra, ema_unet=use_ema_unet)
        self._last_ema_unet = use_ema_unet
        if timesteps is None:
            timesteps = self.sample_timesteps(batched_inputs, bsz)
        cond_inputs = batched_inputs["cond_inputs"]
        cond_emb = batched_inputs["cond_emb"]
        if cond_emb.dim() == 3:  # [B,1,1280] -> [B,1280]  (ldm_diffusers.py:507-508)
            cond_emb = cond_emb[:, 0]
        if cond_inputs.shape[0] == 1 and bsz > 1:
            cond_inputs = cond_inputs.expand(bsz, -1, -1)
        if cond_emb.shape[0] == 1 and bsz > 1:
            cond_emb = cond_emb.expand(bsz, -1)
        self._serial += 1
        if self.variant == "s0" and (stages & _lib.STAGE_UNET):  # the decoded image is the first feature: the decoder stage belongs to the UNet's
            stages |= _lib.STAGE_DEC
        if out is None and 0 < bsz <= eng.graph_max_batch and stages == eng.stage_all and tuple(images.shape[1:]) == (3, 512, 512):
            return eng.extract_graphed(images.float(), cond_inputs, cond_emb, timesteps, self.shared_noise, ema=ema_projections,
                                       stages=stages, want_taps=want_taps, want_latents=want_latents, want_final=want_final,
                                       out_dtype=out_dtype)
        return eng.extract(images, cond_inputs, cond_emb, timesteps, self.shared_noise, ema=ema_projections, stages=stages,
                           want_taps=want_taps, want_latents=want_latents, out=out, want_final=want_final, out_dtype=out_dtype)

    def forward(self, batched_inputs, input_modal, **kwargs):
        """Reference semantics: returns ``[enc_tap, unet_tap16, unet_tap32, unet_tap64]`` as NCHW fp32 tensors
        (order of ldm_diffusers.py:217: encoder features, then unet features in up-path order 16, 32, 64)."""
        res = self.run(batched_inputs, input_modal, stages=_lib.STAGE_VAE | _lib.STAGE_UNET, want_taps=True, **kwargs)
        enc, t64, t32, t16 = res["taps"]  # s0 variant: `enc` is decoder_output (ldm_diffusers.py:199)
        taps = FeatureTaps([enc, t16, t32, t64])
        taps.token = (id(self), self._serial)
        if kwargs.get("return_unet_final_output"):  # ldm_diffusers.py:211-215
            return taps, {"before_vae.decoder": res["unet_sample"], "after_vae.decoder": res["decoded"]}
        return taps


def vae_encoder(vae, images, encoder_block_indices=()):
    """Module-level ``vae_encoder`` of the reference (``ldm_diffusers.py:283-311``), which the meta-arch imports to encode colour
    targets for the ``vae_decoder_loss`` (``mtmadise.py:15,254,345,398,463``): ``images`` [B,3,512,512] already in [-1,1] ->
    ``(latents [B,4,64,64] = mean * 0.18215, features)``.  ``vae`` is the ``LdmDiffusers.vae`` holder; the VAE stage of the engine runs."""
    owner = getattr(vae, "_owner", None)
    ldm = owner() if owner is not None else None
    if ldm is None:
        raise _lib.MadmError("vae_encoder: `vae` must be the .vae of a madm_b200.ldm.LdmDiffusers")
    idx = list(encoder_block_indices)
    if idx not in ([], [5]) or (idx == [5] and ldm.variant != "base"):
        raise NotImplementedError("vae_encoder: encoder_block_indices must be [] (or [5] in the base configuration)")
    bsz = images.shape[0]
    eng = ldm.prepare(ldm._bound_extra)
    dev = ldm.device
    res = eng.extract(images, torch.zeros(bsz, 77, 768, device=dev), torch.zeros(bsz, 1280, device=dev),
                      torch.zeros(bsz, dtype=torch.int64, device=dev), ldm.shared_noise, stages=_lib.STAGE_VAE, want_latents=True,
                      want_taps=bool(idx), img_normalised=True)
    ldm._serial += 1
    return res["latents"], ([res["taps"][0]] if idx else [])


class ClipFeatureProject(nn.Module):
    """ldm_base.py:632-717 for ``input_prefix=False`` (``clip_state='no'``)."""

    def __init__(self, learnable_cond_prompt=False, prompt_in_features=None, prompt_out_features=None, prompt_seq_len=None,
                 learnable_cond_time=False, time_in_features=None, time_out_features=None, time_seq_len=None,
                 time_alpha_cond_size=None, input_prefix=False, without_prompt_alpha=True, multi_layer_prompt=False,
                 init_uncond_prompt=False, uncond_prompt=None):
        super().__init__()
        if input_prefix or multi_layer_prompt or init_uncond_prompt:
            raise NotImplementedError("clip prefix / multi-layer / uncond-initialised prompts are outside the shipped config")
        if not learnable_cond_time:
            raise NotImplementedError("learnable_cond_time=False (no cond_emb; the reference then passes res_time_embedding=None) is "
                                      "outside the shipped config (mtmadise_multi_lora.py:14-41 sets it True)")
        if learnable_cond_prompt and prompt_seq_len != 77:
            raise NotImplementedError("prompt_seq_len != 77 (the reference interpolates uncond_prompt, ldm_base.py:700-706) is outside "
                                      "the shipped config")
        self.learnable_cond_prompt = learnable_cond_prompt
        self.learnable_cond_time = learnable_cond_time
        self.input_prefix = input_prefix
        self.without_prompt_alpha = without_prompt_alpha
        if self.learnable_cond_prompt:
            pe = torch.zeros(1, prompt_seq_len, prompt_out_features)
            nn.init.trunc_normal_(pe, std=0.02, a=-2.0, b=2.0)
            self.prompt_embed = nn.Parameter(pe)
            if not self.without_prompt_alpha:
                shape = [1, prompt_seq_len, prompt_out_features]
                self.alpha_cond_prompt = nn.Parameter(torch.rand(shape))
                self.alpha_uncond_prompt = nn.Parameter(torch.rand(shape))
        if self.learnable_cond_time:
            self.alpha_cond_time = nn.Parameter(torch.zeros(time_alpha_cond_size))
            te = torch.zeros(1, time_seq_len, time_out_features)
            nn.init.trunc_normal_(te, std=0.02, a=-2.0, b=2.0)
            self.time_embed = nn.Parameter(te)

    def get_cond_prompt(self, uncond_prompt, prefix=None):
        if not self.learnable_cond_prompt:
            return uncond_prompt
        if self.without_prompt_alpha:
            return self.prompt_embed
        return torch.tanh(self.alpha_uncond_prompt) * uncond_prompt + torch.tanh(self.alpha_cond_prompt) * self.prompt_embed

    def get_cond_time(self, prefix=None):
        return torch.tanh(self.alpha_cond_time) * self.time_embed if self.learnable_cond_time else None

    def forward(self, uncond_prompt, prefix=None):
        return self.get_cond_prompt(uncond_prompt, prefix), self.get_cond_time(prefix)


class BasePromptTimeGenerator(nn.Module):
    cross_attention_out_dim = [320, 320, 640, 640, 1280, 1280, 1280, 1280, 1280, 1280, 640, 640, 640, 320, 320, 320]

    def __init__(self, learnable_cond_prompt=True, learnable_cond_time=True, same_cond_params=False,
                 detach_prompt_for_mixed_data=False, clip_state="no", num_timesteps=1, clip_model_name="", ldm_extractor=None,
                 without_prompt_alpha=False, multi_layer_prompt=False, mix_source_target_prompt=False, init_uncond_prompt=False,
                 mask_prompt_ratio=False, detach_mask_prompt=False, prompt_perturbation=False, rand_prompt_scale=None, **kwargs):
        super().__init__()
        assert clip_state in {"no", "learnable_clip", "no_learnable_clip"}
        if clip_state != "no":
            raise NotImplementedError("clip_state != 'no' (OpenCLIP prefix) is out of scope; the shipped config sets 'no'")
        if ldm_extractor is None:
            raise NotImplementedError("the legacy CompVis LdmExtractor path is out of scope; pass ldm_extractor=LdmDiffusers(...)")
        self.learnable_cond_prompt = learnable_cond_prompt
        self.learnable_cond_time = learnable_cond_time
        self.same_cond_params = same_cond_params
        self.detach_prompt_for_mixed_data = detach_prompt_for_mixed_data
        self.clip_state = clip_state
        self.multi_layer_prompt = multi_layer_prompt
        self.without_prompt_alpha = without_prompt_alpha
        self.mix_source_target_prompt = mix_source_target_prompt
        self.init_uncond_prompt = init_uncond_prompt
        self.mask_prompt_ratio = mask_prompt_ratio
        self.detach_mask_prompt = detach_mask_prompt
        assert not (self.detach_mask_prompt and (not self.mask_prompt_ratio))
        self.prompt_perturbation = prompt_perturbation
        self.rand_prompt_scale = rand_prompt_scale
        self.ldm_extractor = ldm_extractor
        self.text_embed_shape = ldm_extractor.text_embed_shape
        t_out = ldm_extractor.unet_time_embed_out_features
        seq = kwargs.get("prompt_seq_len", self.text_embed_shape[0])

        def mk():
            return ClipFeatureProject(
                learnable_cond_prompt=learnable_cond_prompt, prompt_in_features=None, prompt_out_features=self.text_embed_shape[1],
                prompt_seq_len=seq, without_prompt_alpha=without_prompt_alpha, multi_layer_prompt=multi_layer_prompt,
                init_uncond_prompt=init_uncond_prompt, uncond_prompt=None, learnable_cond_time=learnable_cond_time,
                time_in_features=None, time_out_features=t_out, time_seq_len=num_timesteps, time_alpha_cond_size=t_out,
                input_prefix=False).to(ldm_extractor.device)

        self.clip_project_rgb = mk()
        self.clip_project_others = self.clip_project_rgb if same_cond_params else mk()

    @property
    def uncond_inputs(self):
        return self.ldm_extractor.uncond_inputs

    def conditioning(self, batched_inputs, input_modal, ema_forward=False, timestep=None):
        """ldm_base.py:832-917: fills cond_inputs / cond_emb / timestep into ``batched_inputs``."""
        assert input_modal in {"rgb", "others", "mixed", "masked_prompt", "prompt_perturbation", "rand_prompt"}
        image = batched_inputs["img"]
        # The deterministic modes depend on a handful of small parameters only (batch-invariant): the batched tensors are cached
        # per (mode, ema, batch) and re-derived when a parameter's version counter moves, so steady-state inference launches no
        # PyTorch kernels for the conditioning at all.
        cache_key = None
        if not torch.is_grad_enabled() and (input_modal in ("rgb", "others") or (input_modal == "mixed" and self.mix_source_target_prompt)):
            mods = [self.clip_project_rgb] if input_modal == "rgb" else (
                [self.clip_project_rgb, self.clip_project_others] if input_modal == "mixed"
                else [self.ema_clip_project_others if ema_forward else self.clip_project_others])
            vers = tuple((id(t), t._version, t.data_ptr()) for m in mods for t in m.parameters()) + (
                (id(self.uncond_inputs), self.uncond_inputs._version),)
            cache_key = (input_modal, bool(ema_forward), int(image.shape[0]), str(image.device))
            hit = getattr(self, "_cond_cache", {}).get(cache_key)
            if hit is not None and hit[0] == vers:
                if timestep is not None:
                    batched_inputs["timestep"] = timestep
                batched_inputs["cond_inputs"], batched_inputs["cond_emb"] = hit[1], hit[2]
                return batched_inputs
        if input_modal == "rgb":
            assert ema_forward is False
            ci, ce = self.clip_project_rgb(self.uncond_inputs, None)
        elif input_modal == "mixed" and self.mix_source_target_prompt:
            s_ci, s_ce = self.clip_project_rgb(self.uncond_inputs, None)
            t_ci, t_ce = self.clip_project_others(self.uncond_inputs, None)
            ci, ce = (s_ci + t_ci) / 2, (s_ce + t_ce) / 2
        else:
            proj = self.ema_clip_project_others if ema_forward else self.clip_project_others
            ci, ce = proj(self.uncond_inputs, None)
        if input_modal == "mixed" and self.detach_prompt_for_mixed_data:
            ci = ci.detach()
        if input_modal == "masked_prompt" and self.mask_prompt_ratio:
            ci = self.mask_prompt(ci).detach() if self.detach_mask_prompt else self.mask_prompt(ci)
        elif input_modal == "prompt_perturbation" and self.prompt_perturbation:
            ci = (ci + torch.randn(ci.shape, device=ci.device) * self.prompt_perturbation).detach()
        elif input_modal == "rand_prompt":
            ci = torch.rand_like(ci) * self.rand_prompt_scale
        if timestep is not None:
            batched_inputs["timestep"] = timestep
        if image.shape[0] != 1:
            ci = torch.repeat_interleave(ci, repeats=image.shape[0], dim=0)
            ce = torch.repeat_interleave(ce, repeats=image.shape[0], dim=0)
        batched_inputs["cond_inputs"], batched_inputs["cond_emb"] = ci, ce
        if cache_key is not None:
            if not hasattr(self, "_cond_cache"):
                object.__setattr__(self, "_cond_cache", {})
            self._cond_cache[cache_key] = (vers, ci.detach(), ce.detach())
        return batched_inputs

    def forward(self, batched_inputs, input_modal, ema_forward=False, timestep=None, return_unet_feats=False, **kwargs):
        assert not return_unet_feats
        with torch.no_grad():
            batched_inputs = self.conditioning(batched_inputs, input_modal, ema_forward, timestep)
        return self.ldm_extractor(batched_inputs, input_modal, ema_forward=ema_forward, **kwargs)

    def mask_prompt(self, prompt):  # ldm_base.py:926-938
        assert prompt.dim() == 3
        mask = (torch.rand((prompt.shape[0], prompt.shape[1], 1), device=prompt.device) > self.mask_prompt_ratio).float()
        return prompt * mask

    feature_size = property(lambda self: self.ldm_extractor.feature_size)
    feature_dims = property(lambda self: self.ldm_extractor.feature_dims)
    feature_strides = property(lambda self: self.ldm_extractor.feature_strides)
    num_groups = property(lambda self: self.ldm_extractor.num_groups)
    grouped_indices = property(lambda self: self.ldm_extractor.grouped_indices)

    def extra_repr(self):
        return f"learnable_time_embed={self.learnable_cond_time}"

    def set_requires_grad(self, requires_grad):
        for p in self.ldm_extractor.unet.parameters():
            p.requires_grad = requires_grad

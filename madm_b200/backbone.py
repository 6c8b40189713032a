"""Drop-in mirror of the reference backbone boundary
(``modeling/backbone/feature_extractor.py:20-396``): same constructor kwargs as
``config_files/common/models/mtmadise_multi_lora.py:14-41``, same ``forward(img, input_modal, ema_forward,
timestep, **kwargs)`` signature, same ``{'output_features': {'s2','s3','s4','s5'}}`` result and the same
state_dict key names, so only the LazyCall ``_target_`` changes.

The forward pass is one ``madm_extract`` call: VAE encode -> q-sample -> UNet with taps -> GN-bottleneck
projections, all in the CUDA engine.  There is no PyTorch/eager fallback.
"""
import logging
from collections import OrderedDict
from typing import List, Tuple, Union

import torch
import torch.nn as nn

from . import _lib
from .ldm import BasePromptTimeGenerator, FeatureTaps, LdmDiffusers
from .sd14_params import BottleneckParams

logger = logging.getLogger(__name__)

try:  # inside MADM (detectron2 installed) the class must be a detectron2 Backbone
    from detectron2.modeling.backbone import Backbone as _BackboneBase  # type: ignore
except Exception:
    _BackboneBase = nn.Module


def make_projection(cin: int, cout: int, bottleneck_channels: int, num_blocks: int, device) -> nn.Sequential:
    """``nn.Sequential(*ResNet.make_stage(BottleneckBlock, num_blocks, in, bottleneck, out, norm='GN'))`` as holders."""
    if num_blocks != 1:
        raise NotImplementedError("num_res_blocks != 1 is outside the shipped config")
    return nn.Sequential(BottleneckParams(cin, cout, bottleneck_channels, device=device))


class FeatureExtractorBackbone(_BackboneBase):
    """feature_extractor.py:20-284 (constructor bookkeeping, preprocess, single / slide forward)."""

    def __init__(self, feature_extractor, out_features: List[str], backbone_in_size: Union[int, Tuple[int]] = (512, 512),
                 min_stride: int = 4, max_stride: int = 32, projection_dim=512, num_res_blocks: int = 1,
                 use_checkpoint: bool = False, slide_training: bool = False, slide_inference: bool = False):
        super().__init__()
        self.feature_extractor = feature_extractor
        self.use_checkpoint = use_checkpoint
        self._slide_inference = slide_inference
        self.backbone_in_size = tuple(backbone_in_size) if not isinstance(backbone_in_size, int) else (backbone_in_size,) * 2
        if self.backbone_in_size != (512, 512):
            raise NotImplementedError("backbone_in_size must be (512, 512): the SD-1.4 latent grid is fixed at 64x64")
        self._slide_training = slide_training
        if self._slide_training:
            assert self._slide_inference, "slide training must be used with slide inference"
        self.y1_y2_x1_x2 = [(0, 512, 0, 512), (0, 512, 256, 768), (0, 512, 512, 1024)] if slide_inference else None  # :75
        self.min_stride = min_stride
        self.max_stride = max_stride
        self._out_feature_channels = {}
        self._out_feature_strides = {}
        self._out_features = list(out_features)

    @property
    def size_divisibility(self) -> int:  # :127-129
        return 64

    def ignored_state_dict(self, destination=None, prefix=""):  # :131-138
        if destination is None:
            destination = OrderedDict()
            destination._metadata = OrderedDict()
        for name, module in self._modules.items():
            if module is not None and hasattr(module, "ignored_state_dict"):
                module.ignored_state_dict(destination, prefix + name + ".")
        return destination

    def preprocess_image(self, img):  # :140-146
        """T.Resize(backbone_in_size, BILINEAR) unless sliding (with the reference's pinned torchvision 0.16.1 a tensor input is NOT
        antialiased), then ImageList.from_tensors' zero padding to a multiple of 64: one kernel; identity (no launch) at 512 x 512."""
        resize = not self._slide_inference and tuple(img.shape[-2:]) != self.backbone_in_size
        h, w = self.backbone_in_size if resize else img.shape[-2:]
        if not resize and h % self.size_divisibility == 0 and w % self.size_divisibility == 0:
            return img
        from . import ops
        return ops.preprocess_image(img, self.backbone_in_size if resize else None, self.size_divisibility)

    def checkpoint_forward_features(self, features, input_image_size, ema_forward=False):  # :148-154
        return self.forward_features(features, input_image_size, ema_forward)

    def slide_windows(self, h_img: int, w_img: int, crop: int = 512, stride: int = 256):
        """512x512 windows at stride 256 on both axes.  On a 512x1024 input this is exactly the
        reference's three hard-coded windows (:75); the reference's own general path is inconsistent."""
        if h_img < crop or w_img < crop or (h_img - crop) % stride or (w_img - crop) % stride:
            raise ValueError(f"slide_forward needs H, W = 512 + k*256, got {(h_img, w_img)}")
        return [(y, y + crop, x, x + crop) for y in range(0, h_img - crop + 1, stride) for x in range(0, w_img - crop + 1, stride)]

    def forward(self, img, input_modal="rgb", ema_forward=False, timestep=None, **kwargs):  # :280-284
        if (self.training and not self._slide_training) or not self._slide_inference:
            return self.single_forward(img, input_modal, ema_forward, timestep, **kwargs)
        return self.slide_forward(img, input_modal, ema_forward, timestep, **kwargs)


class AttentionFeatureExtractorBackbone(FeatureExtractorBackbone):
    """feature_extractor.py:287-396."""

    def __init__(self, attention_features_res, feature_dims, attention_features_location, target_attention_loss=False,
                 attention_select_index=None, feature_extractor=None, out_features: List[str] = None,
                 backbone_in_size: Union[int, Tuple[int]] = (512, 512), min_stride: int = 4, max_stride: int = 32,
                 projection_dim: List[int] = [512, 512, 512, 512], bottleneck_channels: int = 512 // 4, num_res_blocks: int = 1,
                 use_checkpoint: bool = False, slide_training: bool = False, slide_inference: bool = False, crop_batch: int = 16,
                 feature_dtype=torch.float32):
        super().__init__(feature_extractor, out_features, backbone_in_size, min_stride, max_stride, projection_dim, num_res_blocks,
                         use_checkpoint, slide_training, slide_inference)
        # extension (not a reference kwarg): torch.float16 makes inference return fp16 feature maps (base variant) -- half the bytes for
        # callers that download the feature dict; the reference's maps are fp32, which stays the default
        self.feature_dtype = feature_dtype
        self.attention_features_res = attention_features_res
        self.feature_dims = list(feature_dims)
        self.attention_features_location = attention_features_location
        self.target_attention_loss = target_attention_loss
        self.attention_select_index = attention_select_index
        self.crop_batch = crop_batch
        variant = getattr(feature_extractor.ldm_extractor, "variant", "base")
        want = dict(base=(["s2", "s3", "s4", "s5"], [512, 320, 640, 1280], [512] * 4),
                    s0=(["s0", "s3", "s4", "s5"], [3, 320, 640, 1280], [128, 512, 512, 512]))[variant]
        if (list(out_features), self.feature_dims, list(projection_dim)) != want or bottleneck_channels != 128:
            raise NotImplementedError(
                "madm_b200 implements the shipped projection configs: out_features s2..s5 / feature_dims [512,320,640,1280] / "
                "projection_dim [512]*4 (mtmadise_multi_lora.py:14-41), or with ldm_extractor.vae_decoder_loss=True out_features "
                "s0,s3,s4,s5 / feature_dims [3,320,640,1280] / projection_dim [128,512,512,512] "
                "(mtmadise_cityscapes_rgb_to_depth_11.py:47-55); bottleneck 128")
        self.variant = variant
        device = feature_extractor.ldm_extractor.device
        self.feature_projections = nn.ModuleList(
            [make_projection(fd, projection_dim[i], bottleneck_channels, num_res_blocks, device) for i, fd in enumerate(self.feature_dims)])
        self.register_load_state_dict_post_hook(lambda module, incompatible: object.__setattr__(module, "_proj_cache", None))
        self._out_feature_strides = {s: 2 ** int(s[1]) for s in out_features}  # :361-364
        self._out_feature_channels = {s: projection_dim[i] for i, s in enumerate(out_features)}
        self._out_features = list(self._out_feature_strides.keys())

    # ------------------------------------------------------------------ engine plumbing
    def _projection_tensors(self) -> List[Tuple[str, torch.Tensor]]:
        ema = getattr(self, "ema_feature_projections", None)  # set by CMDISE._inti_ema_weights (cmdise.py:308)
        key = (id(self.feature_projections), id(ema))
        hit = getattr(self, "_proj_cache", None)
        if hit is not None and hit[0] == key:
            return hit[1]
        out = [("feature_projections." + n, p) for n, p in self.feature_projections.named_parameters()]
        if ema is not None:
            out += [("ema_feature_projections." + n, p) for n, p in ema.named_parameters()]
        object.__setattr__(self, "_proj_cache", (key, out))
        return out

    def _grad_inputs(self, input_modal, ema_forward) -> List[Tuple[str, torch.Tensor]]:
        """Parameters this call's result depends on that currently require grad (empty under no_grad)."""
        if not torch.is_grad_enabled():
            return []
        gen: BasePromptTimeGenerator = self.feature_extractor
        ldm: LdmDiffusers = gen.ldm_extractor
        use_ema_unet = bool(ema_forward and hasattr(ldm, "ema_unet"))
        # the big lists (UNet, projections) come from the caches the engine binding uses; names already carry their prefixes
        cands = [(n, p) for n, p in ldm.named_engine_tensors(use_ema_unet) if ".unet." in n]
        cands += [(("feature_projections." + n[len("ema_feature_projections."):]) if n.startswith("ema_") else n, p)
                  for n, p in self._projection_tensors() if n.startswith("ema_feature_projections.") == bool(ema_forward)]
        if input_modal in ("rgb", "mixed"):
            cands += [("feature_extractor.clip_project_rgb." + n, p) for n, p in gen.clip_project_rgb.named_parameters()]
        if input_modal != "rgb":
            others = gen.ema_clip_project_others if ema_forward else gen.clip_project_others
            cands += [("feature_extractor.clip_project_others." + n, p) for n, p in others.named_parameters()]
        seen, out = set(), []
        for n, p in cands:
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                out.append((n, p))
        return out

    def _extract(self, img, input_modal, ema_forward, timestep, want_taps=False, timesteps=None, **kwargs):
        gen: BasePromptTimeGenerator = self.feature_extractor
        grad_inputs = self._grad_inputs(input_modal, ema_forward)
        if grad_inputs:
            # Called under grad with trainable parameters (MTMADISE.forward's student passes, mtmadise.py:240-302): the result must carry
            # a grad_fn.  The training path (madm_b200/train.py: autograd.Function around madm_extract / madm_backward) provides it for
            # the LoRA training step's trainable set and raises for anything else -- never a silent grad-free tensor.
            from . import train
            return train.extract_with_grad(self, img, input_modal, ema_forward, timestep, grad_inputs, want_taps=want_taps,
                                           timesteps=timesteps, **kwargs)
        batched = dict(img=img)
        with torch.no_grad():
            gen.conditioning(batched, input_modal, ema_forward, timestep)
        if ema_forward and not hasattr(self, "ema_feature_projections"):
            raise AttributeError("ema_forward=True needs backbone.ema_feature_projections (CMDISE._inti_ema_weights)")
        return gen.ldm_extractor.run(batched, input_modal, stages=_lib.STAGE_ALL, ema_projections=bool(ema_forward),
                                     extra=self._projection_tensors(), want_taps=want_taps, timesteps=timesteps,
                                     ema_forward=ema_forward, out_dtype=getattr(self, "feature_dtype", torch.float32), **kwargs)

    # ------------------------------------------------------------------ reference surface
    def single_forward(self, img, input_modal="rgb", ema_forward=False, timestep=None, **kwargs):  # :156-170
        img = self.preprocess_image(img)
        if tuple(img.shape[-2:]) != (512, 512):
            raise ValueError(f"single_forward expects 512x512 after preprocessing, got {tuple(img.shape[-2:])}")
        res = self._extract(img, input_modal, ema_forward, timestep, **kwargs)
        feats = {"output_features": dict(zip(self._out_features, res["features"]))}
        if "return_unet_final_output" in kwargs:  # :164-166
            return feats, {"before_vae.decoder": res["unet_sample"], "after_vae.decoder": res["decoded"]}
        return feats

    def forward_features(self, features, input_image_size=None, ema_forward=False):  # :367-396
        """Projection stage on taps produced by ``self.feature_extractor(...)`` (the reference's two-step use)."""
        if not isinstance(features, FeatureTaps) or features.token is None:
            raise NotImplementedError("forward_features needs the FeatureTaps returned by this backbone's feature_extractor")
        if torch.is_grad_enabled() and any(p.requires_grad for p in (self.ema_feature_projections if ema_forward else self.feature_projections).parameters()):
            raise NotImplementedError("forward_features under torch.enable_grad() with trainable projections: the two-step use has no "
                                      "backward path; call the backbone's forward (training path) or wrap inference in torch.no_grad()")
        ldm: LdmDiffusers = self.feature_extractor.ldm_extractor
        if features.token != (id(ldm), ldm._serial):
            raise RuntimeError("stale FeatureTaps: another forward ran since these taps were produced")
        b = features[0].shape[0]
        eng = ldm.prepare(self._projection_tensors(), ema_unet=ldm._last_ema_unet)  # the context whose workspace holds these taps
        dummy = torch.zeros(b, dtype=torch.int64, device=ldm.device)
        res = eng.extract(None, torch.zeros(b, 77, 768, device=ldm.device), torch.zeros(b, 1280, device=ldm.device), dummy,
                          ldm.shared_noise, ema=bool(ema_forward), stages=_lib.STAGE_PROJ, B=b)
        return {"output_features": dict(zip(self._out_features, res["features"]))}

    def slide_forward(self, img, input_modal="rgb", ema_forward=False, timestep=None, **kwargs):  # :199-278
        if "return_unet_final_output" in kwargs:
            raise NotImplementedError("return_unet_final_output with slide inference (the reference's slide_forward drops it too)")
        b, _, h_img, w_img = img.shape
        wins = self.slide_windows(h_img, w_img)
        from . import ops
        # crops are the batch dimension of the engine: all windows of `crop_batch // len(wins)` images per call; the accumulate /
        # divide-by-count of feature_extractor.py:254-275 is one gather kernel per feature map (window order = the reference's += order)
        per_call = max(1, self.crop_batch // len(wins))
        parts = {k: [] for k in self._out_features}
        for i0 in range(0, b, per_call):
            i1 = min(b, i0 + per_call)
            crops = torch.cat([img[i0:i1, :, y1:y2, x1:x2] for (y1, y2, x1, x2) in wins], dim=0)
            feats = self._extract(crops, input_modal, ema_forward, timestep, **kwargs)["features"]
            for k, f in zip(self._out_features, feats):
                s = self._out_feature_strides[k]
                f = f if f.dtype == torch.float32 else f.float()  # (feature_dtype=float16: the merge accumulates in fp32)
                parts[k].append(ops.slide_merge(f, len(wins), [(y1 // s, x1 // s) for (y1, _, x1, _) in wins], h_img // s, w_img // s))
        outs = {k: (v[0] if len(v) == 1 else torch.cat(v, dim=0)) for k, v in parts.items()}
        return {"output_features": outs}

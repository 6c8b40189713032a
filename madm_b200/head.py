"""Host-side mirror of MADM's ``DAFormerHead`` (reference ``modeling/sem_seg_head/daformer_head.py:536-749``) for the shipped
decoder configuration (``config_files/common/models/mtmadise_multi_lora.py:42-63``: MLP embeds, depthwise-separable ASPP fusion
with dilations (1, 6, 12, 18), BatchNorm + ReLU) — SURVEY §8 row f-2, the consumer right after the feature-extraction path.

Same constructor keywords, ``forward(input_dict)`` contract and ``state_dict`` key names (mmcv ``ConvModule`` = ``.conv`` / ``.bn``,
``DepthwiseSeparableConvModule`` = ``.depthwise_conv`` / ``.pointwise_conv``) as the reference, so MADM checkpoints load with
``load_state_dict``.  The modules below only OWN parameters; all arithmetic runs in ``libmadm_b200.so`` (``MADM_STAGE_HEAD`` of
``madm_extract``: tcgen05 GEMMs with the eval-mode BatchNorm folded into the packed weights, bilinear / depthwise kernels).
Inference only (BatchNorm in eval mode, dropout = identity); there is no eager fallback.
"""
import math
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .engine import Engine

_UNSUPPORTED = "madm_b200.head.DAFormerHead supports the shipped decoder configuration only ({}); SURVEY §8 f-1/f-2 variants are next"


class _Params(nn.Module):
    """Parameter / buffer holder: never called."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter holder: the computation runs in libmadm_b200.so")


class _Conv(_Params):
    def __init__(self, cout, cin_per_group, k, bias=False):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin_per_group, k, k))
        nn.init.kaiming_normal_(self.weight, mode="fan_out", nonlinearity="relu")  # mmcv ConvModule.init_weights
        if bias:
            self.bias = nn.Parameter(torch.zeros(cout))


class _BatchNorm(_Params):
    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class _ConvModule(_Params):  # mmcv ConvModule(conv -> BN -> ReLU), conv has no bias when a norm follows
    def __init__(self, cin, cout, k, groups=1):
        super().__init__()
        self.conv = _Conv(cout, cin // groups, k)
        self.bn = _BatchNorm(cout)


class _DepthwiseSeparable(_Params):
    def __init__(self, cin, cout):
        super().__init__()
        self.depthwise_conv = _ConvModule(cin, cin, 3, groups=cin)
        self.pointwise_conv = _ConvModule(cin, cout, 1)


class _Linear(_Params):
    def __init__(self, cin, cout):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin))
        self.bias = nn.Parameter(torch.empty(cout))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))  # nn.Linear.reset_parameters
        bound = 1.0 / math.sqrt(cin)
        nn.init.uniform_(self.bias, -bound, bound)


class _MLP(_Params):
    def __init__(self, cin, cout):
        super().__init__()
        self.proj = _Linear(cin, cout)


class _ASPPWrapper(_Params):
    def __init__(self, cin, channels, dilations):
        super().__init__()
        self.aspp_modules = nn.ModuleList([_ConvModule(cin, channels, 1) if d == 1 else _DepthwiseSeparable(cin, channels) for d in dilations])
        self.bottleneck = _ConvModule(len(dilations) * channels, channels, 3)


class DAFormerHead(nn.Module):
    def __init__(self, in_channels, in_keys, channels, *, num_classes, dropout_ratio=0.1, conv_cfg=None, norm_cfg=None,
                 act_cfg=dict(type="ReLU"), in_index=-1, input_transform="multiple_select", decoder_params=None, ignore_index=255,
                 align_corners=False, init_cfg=dict(type="Normal", std=0.01, override=dict(name="conv_seg")),
                 concat_attention_to_conv_seg=False, final_fuse_vae_decoder_feat=False, device=None, compute_dtype: str = "fp16"):
        super().__init__()
        in_channels, in_keys = list(in_channels), list(in_keys)
        in_index = list(in_index) if not isinstance(in_index, int) else [in_index]
        dp = dict(decoder_params or {})
        fusion = dict(dp.get("fusion_cfg") or {})
        embed = dict(dp.get("embed_cfg") or {})
        neck = dp.get("embed_neck_cfg")
        neck = embed if neck == "same_as_embed_cfg" else dict(neck or {})
        checks = [
            (not concat_attention_to_conv_seg, "concat_attention_to_conv_seg=False"),
            (not final_fuse_vae_decoder_feat, "final_fuse_vae_decoder_feat=False"),
            (not align_corners, "align_corners=False"),
            (input_transform == "multiple_select", "input_transform='multiple_select'"),
            ((in_channels, in_keys) in (([512] * 4, ["s2", "s3", "s4", "s5"]), ([128, 512, 512, 512], ["s0", "s3", "s4", "s5"]))
             and in_index == [0, 1, 2, 3], "the four 512-channel maps s2..s5, or s0 (128 channels) + s3..s5 "
             "(mtmadise_cityscapes_rgb_to_depth_11.py:51-55)"),
            (embed.get("type") == "mlp" and neck.get("type") == "mlp", "embed_cfg / embed_neck_cfg type 'mlp'"),
            (fusion.get("type") == "aspp" and bool(fusion.get("sep")) and tuple(fusion.get("dilations", ())) == (1, 6, 12, 18)
             and not fusion.get("pool") and not fusion.get("context_cfg"), "fusion_cfg aspp, sep=True, dilations (1,6,12,18), pool=False"),
            ((norm_cfg or {}).get("type") == "BN" and (fusion.get("norm_cfg") or {}).get("type") == "BN", "norm_cfg BN"),
            ((act_cfg or {}).get("type") == "ReLU", "act_cfg ReLU"),
            (isinstance(dp.get("embed_dims"), int) and dp["embed_dims"] % 64 == 0 and channels % 64 == 0 and num_classes <= 32,
             "embed_dims / channels multiples of 64, num_classes <= 32"),
        ]
        for ok, what in checks:
            if not ok:
                raise NotImplementedError(_UNSUPPORTED.format(what))
        self.in_channels, self.in_keys, self.in_index = in_channels, in_keys, in_index
        self.variant = "s0" if in_keys[0] == "s0" else "base"
        self.channels, self.num_classes, self.dropout_ratio = channels, num_classes, dropout_ratio
        self.ignore_index, self.align_corners = ignore_index, align_corners
        E = dp["embed_dims"]
        self.embed_layers = nn.ModuleDict({str(i): _MLP(c, E) for i, c in zip(in_index, in_channels)})
        self.fuse_layer = _ASPPWrapper(E * len(in_index), channels, (1, 6, 12, 18))
        self.conv_seg = _Conv(num_classes, channels, 1, bias=True)
        nn.init.normal_(self.conv_seg.weight, std=0.01)  # init_cfg Normal(std=0.01) override conv_seg (daformer_head.py:553)
        self.compute_dtype = compute_dtype
        self._engine: Optional[Engine] = None
        if device is not None:
            self.to(device)

    # ------------------------------------------------------------------ engine plumbing
    def _named_tensors(self):
        out = []
        for n, t in list(self.named_parameters()) + list(self.named_buffers()):
            if t.dtype == torch.float32:
                out.append(("sem_seg_head." + n, t.detach()))
        return out

    def engine(self) -> Engine:
        dev = self.conv_seg.weight.device
        if self._engine is None or self._engine.device != dev:
            self._engine = Engine(dev, self.compute_dtype, self.variant)  # raises without libmadm_b200.so / an sm_100 device
        return self._engine

    def transfer_input_dict_to_list(self, inputs_dict: Dict[str, torch.Tensor]) -> List[torch.Tensor]:
        inputs_list = [inputs_dict[k] for k in self.in_keys]
        assert len(inputs_list) == len(inputs_dict)  # every backbone feature is used (daformer_head.py:661-671)
        return inputs_list

    def forward(self, input_dict):
        if self.training:
            raise NotImplementedError("madm_b200.head.DAFormerHead is inference-only (eval-mode BatchNorm folded into the weights); "
                                      "call .eval() — the training step is SURVEY §8 row f-3")
        if "cross_attention_feat" in input_dict:
            raise NotImplementedError(_UNSUPPORTED.format("no cross_attention_feat"))
        x = self.transfer_input_dict_to_list(input_dict["output_features"])
        eng = self.engine()
        eng.bind(self._named_tensors())
        eng.ensure_packed(None, 0.0)
        return eng.head(x, self.num_classes)

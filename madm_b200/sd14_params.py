"""Parameter holders for the SD-1.4 UNet / VAE with diffusers (and peft) state_dict key names.

The product never runs these weights through PyTorch: the modules below only *own* the fp32
``nn.Parameter``s under the names MADM checkpoints use (reference
``checkpoint/odise_checkpointer.py:38-111``), so ``load_state_dict``, optimizers, EMA and
``named_parameters()``-based code in the reference keep working, while the forward pass is the CUDA
engine reading the same storage through raw pointers.

LoRA: ``unet.add_adapter(config, name)`` / ``unet.set_adapter(names)`` mirror what MADM calls on the
diffusers UNet (reference ``modeling/meta_arch/mtmadise.py:115-147``); wrapped projections expose
``base_layer`` / ``lora_A.<name>`` / ``lora_B.<name>`` exactly like peft 0.10.0, and
``module._active_adapter`` is the attribute the reference writes per forward.
"""
import math
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

try:  # when peft is installed the reference's isinstance(module, BaseTunerLayer) checks must see our layers
    from peft.tuners.tuners_utils import BaseTunerLayer as _TunerBase  # type: ignore
except Exception:  # peft absent (this image): standalone marker class
    class _TunerBase:  # type: ignore
        pass

LORA_TARGETS = ("to_k", "to_q", "to_v", "to_out.0")
Spec = List[Tuple[str, Tuple[int, ...]]]

_EMPTY_INIT = False


class empty_init:
    """``with empty_init(): ...`` builds holders with uninitialised storage (``torch.empty``: no random-init kernels, no host work)
    for callers that fill every parameter right afterwards with ``load_state_dict`` — the normal MADM workflow (checkpoint load)."""

    def __enter__(self):
        global _EMPTY_INIT
        self._prev, _EMPTY_INIT = _EMPTY_INIT, True
        return self

    def __exit__(self, *exc):
        global _EMPTY_INIT
        _EMPTY_INIT = self._prev
        return False


# ----------------------------------------------------------------------------------------- specs
def _res(p: str, cin: int, cout: int, temb: Optional[int]) -> Spec:
    s = [(f"{p}.norm1.weight", (cin,)), (f"{p}.norm1.bias", (cin,)),
         (f"{p}.conv1.weight", (cout, cin, 3, 3)), (f"{p}.conv1.bias", (cout,))]
    if temb:
        s += [(f"{p}.time_emb_proj.weight", (cout, temb)), (f"{p}.time_emb_proj.bias", (cout,))]
    s += [(f"{p}.norm2.weight", (cout,)), (f"{p}.norm2.bias", (cout,)),
          (f"{p}.conv2.weight", (cout, cout, 3, 3)), (f"{p}.conv2.bias", (cout,))]
    if cin != cout:
        s += [(f"{p}.conv_shortcut.weight", (cout, cin, 1, 1)), (f"{p}.conv_shortcut.bias", (cout,))]
    return s


def _attn(p: str, c: int, kv: int, qkv_bias: bool) -> Spec:
    s: Spec = []
    for n, k in (("to_q", c), ("to_k", kv), ("to_v", kv)):
        s.append((f"{p}.{n}.weight", (c, k)))
        if qkv_bias:
            s.append((f"{p}.{n}.bias", (c,)))
    s += [(f"{p}.to_out.0.weight", (c, c)), (f"{p}.to_out.0.bias", (c,))]
    return s


def _transformer(p: str, c: int, cross: int) -> Spec:
    t = f"{p}.transformer_blocks.0"
    s = [(f"{p}.norm.weight", (c,)), (f"{p}.norm.bias", (c,)), (f"{p}.proj_in.weight", (c, c, 1, 1)), (f"{p}.proj_in.bias", (c,))]
    s += [(f"{t}.norm1.weight", (c,)), (f"{t}.norm1.bias", (c,))] + _attn(f"{t}.attn1", c, c, False)
    s += [(f"{t}.norm2.weight", (c,)), (f"{t}.norm2.bias", (c,))] + _attn(f"{t}.attn2", c, cross, False)
    s += [(f"{t}.norm3.weight", (c,)), (f"{t}.norm3.bias", (c,)),
          (f"{t}.ff.net.0.proj.weight", (8 * c, c)), (f"{t}.ff.net.0.proj.bias", (8 * c,)),
          (f"{t}.ff.net.2.weight", (c, 4 * c)), (f"{t}.ff.net.2.bias", (c,))]
    s += [(f"{p}.proj_out.weight", (c, c, 1, 1)), (f"{p}.proj_out.bias", (c,))]
    return s


def unet_spec(in_channels: int = 4, out_channels: int = 4, cross: int = 768) -> Spec:
    """SD-1.4 ``unet/config.json``: 859,520,964 parameters."""
    ch = (320, 640, 1280, 1280)
    temb = 1280
    s: Spec = [("conv_in.weight", (ch[0], in_channels, 3, 3)), ("conv_in.bias", (ch[0],)),
               ("time_embedding.linear_1.weight", (temb, ch[0])), ("time_embedding.linear_1.bias", (temb,)),
               ("time_embedding.linear_2.weight", (temb, temb)), ("time_embedding.linear_2.bias", (temb,))]
    cout = ch[0]
    for i in range(4):
        cin, cout = cout, ch[i]
        for j in range(2):
            s += _res(f"down_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout, temb)
        if i < 3:
            for j in range(2):
                s += _transformer(f"down_blocks.{i}.attentions.{j}", cout, cross)
            s += [(f"down_blocks.{i}.downsamplers.0.conv.weight", (cout, cout, 3, 3)), (f"down_blocks.{i}.downsamplers.0.conv.bias", (cout,))]
    s += _res("mid_block.resnets.0", 1280, 1280, temb) + _transformer("mid_block.attentions.0", 1280, cross) + \
        _res("mid_block.resnets.1", 1280, 1280, temb)
    rev = ch[::-1]
    cout = rev[0]
    for i in range(4):
        prev, cout = cout, rev[i]
        cin = rev[min(i + 1, 3)]
        for j in range(3):
            skip = cin if j == 2 else cout
            rin = prev if j == 0 else cout
            s += _res(f"up_blocks.{i}.resnets.{j}", rin + skip, cout, temb)
        if i > 0:
            for j in range(3):
                s += _transformer(f"up_blocks.{i}.attentions.{j}", cout, cross)
        if i < 3:
            s += [(f"up_blocks.{i}.upsamplers.0.conv.weight", (cout, cout, 3, 3)), (f"up_blocks.{i}.upsamplers.0.conv.bias", (cout,))]
    s += [("conv_norm_out.weight", (ch[0],)), ("conv_norm_out.bias", (ch[0],)),
          ("conv_out.weight", (out_channels, ch[0], 3, 3)), ("conv_out.bias", (out_channels,))]
    return s


def vae_decoder_spec() -> Spec:
    """Decoder half of SD-1.4 ``vae/config.json`` (49,490,179 parameters): conv_in 4->512, mid block, four up blocks of three
    ResBlocks with output channels (512, 512, 256, 128) and nearest-2x + conv between, GN + conv 128->3."""
    s: Spec = [("decoder.conv_in.weight", (512, 4, 3, 3)), ("decoder.conv_in.bias", (512,))]
    s += _res("decoder.mid_block.resnets.0", 512, 512, None)
    a = "decoder.mid_block.attentions.0"
    s += [(f"{a}.group_norm.weight", (512,)), (f"{a}.group_norm.bias", (512,))] + _attn(a, 512, 512, True)
    s += _res("decoder.mid_block.resnets.1", 512, 512, None)
    cout = 512
    for i, c in enumerate((512, 512, 256, 128)):
        cin, cout = cout, c
        for j in range(3):
            s += _res(f"decoder.up_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout, None)
        if i < 3:
            s += [(f"decoder.up_blocks.{i}.upsamplers.0.conv.weight", (cout, cout, 3, 3)),
                  (f"decoder.up_blocks.{i}.upsamplers.0.conv.bias", (cout,))]
    s += [("decoder.conv_norm_out.weight", (128,)), ("decoder.conv_norm_out.bias", (128,)),
          ("decoder.conv_out.weight", (3, 128, 3, 3)), ("decoder.conv_out.bias", (3,))]
    return s


def vae_spec() -> Spec:
    """SD-1.4 ``vae/config.json`` encoder half + quant / post_quant convs."""
    bo = (128, 256, 512, 512)
    s: Spec = [("encoder.conv_in.weight", (bo[0], 3, 3, 3)), ("encoder.conv_in.bias", (bo[0],))]
    cout = bo[0]
    for i in range(4):
        cin, cout = cout, bo[i]
        for j in range(2):
            s += _res(f"encoder.down_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout, None)
        if i < 3:
            s += [(f"encoder.down_blocks.{i}.downsamplers.0.conv.weight", (cout, cout, 3, 3)),
                  (f"encoder.down_blocks.{i}.downsamplers.0.conv.bias", (cout,))]
    s += _res("encoder.mid_block.resnets.0", 512, 512, None)
    a = "encoder.mid_block.attentions.0"
    s += [(f"{a}.group_norm.weight", (512,)), (f"{a}.group_norm.bias", (512,))] + _attn(a, 512, 512, True)
    s += _res("encoder.mid_block.resnets.1", 512, 512, None)
    s += [("encoder.conv_norm_out.weight", (512,)), ("encoder.conv_norm_out.bias", (512,)),
          ("encoder.conv_out.weight", (8, 512, 3, 3)), ("encoder.conv_out.bias", (8,)),
          ("quant_conv.weight", (8, 8, 1, 1)), ("quant_conv.bias", (8,)),
          ("post_quant_conv.weight", (4, 4, 1, 1)), ("post_quant_conv.bias", (4,))]
    return s


def bottleneck_spec(cin: int, cout: int, mid: int) -> Spec:
    """detectron2 BottleneckBlock(norm='GN') keys."""
    s: Spec = []
    if cin != cout:
        s += [("shortcut.weight", (cout, cin, 1, 1)), ("shortcut.norm.weight", (cout,)), ("shortcut.norm.bias", (cout,))]
    s += [("conv1.weight", (mid, cin, 1, 1)), ("conv1.norm.weight", (mid,)), ("conv1.norm.bias", (mid,)),
          ("conv2.weight", (mid, mid, 3, 3)), ("conv2.norm.weight", (mid,)), ("conv2.norm.bias", (mid,)),
          ("conv3.weight", (cout, mid, 1, 1)), ("conv3.norm.weight", (cout,)), ("conv3.norm.bias", (cout,))]
    return s


# ----------------------------------------------------------------------------------------- holder tree
class ParamNode(nn.Module):
    """A container node of the holder tree.  It has no forward: compute happens in the CUDA engine."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("madm_b200 parameter holders have no PyTorch forward (no CPU / eager fallback); "
                           "call the backbone, which runs the CUDA engine")


def _default_init(name: str, shape: Sequence[int], kind: str, gen: Optional[torch.Generator], device) -> torch.Tensor:
    leaf = name.rsplit(".", 1)[-1]
    parent = name.rsplit(".", 2)[-2] if name.count(".") >= 1 else ""
    is_norm = parent.startswith("norm") or parent in ("group_norm", "conv_norm_out") or (kind == "d2" and parent == "norm")
    if is_norm:
        return torch.ones(shape, device=device) if leaf == "weight" else torch.zeros(shape, device=device)
    if kind == "d2":  # c2_msra_fill: kaiming_normal_(fan_out, relu)
        fan_out = shape[0] * int(math.prod(shape[2:]))
        return torch.randn(shape, generator=gen, device=device) * math.sqrt(2.0 / fan_out)
    # nn.Conv2d / nn.Linear default: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias
    return None  # filled by caller (needs the sibling weight's fan_in for biases)


def build_tree(spec: Spec, kind: str = "torch", device=None, seed: Optional[int] = None) -> ParamNode:
    root = ParamNode()
    gen = None
    if seed is not None:
        gen = torch.Generator(device=device if device is not None else "cpu").manual_seed(seed)
    fan_in: Dict[str, int] = {}
    for name, shape in spec:
        if name.endswith(".weight") and len(shape) >= 2:
            fan_in[name[:-7]] = int(math.prod(shape[1:]))
    for name, shape in spec:
        parts = name.split(".")
        node = root
        for p in parts[:-1]:
            if p not in node._modules:
                node.add_module(p, ParamNode())
            node = node._modules[p]
        t = torch.empty(shape, device=device) if _EMPTY_INIT else _default_init(name, shape, kind, gen, device)
        if t is None:
            bound = 1.0 / math.sqrt(fan_in.get(name.rsplit(".", 1)[0], max(1, shape[0])))
            t = (torch.rand(shape, generator=gen, device=device) * 2.0 - 1.0) * bound
        node.register_parameter(parts[-1], nn.Parameter(t))
    return root


# ----------------------------------------------------------------------------------------- LoRA
class LoraLinearParams(ParamNode, _TunerBase):
    """peft LoRA ``Linear`` as a parameter holder: base_layer + lora_A/lora_B per adapter."""

    def __init__(self, base: ParamNode):
        nn.Module.__init__(self)
        self.base_layer = base
        self.lora_A = nn.ModuleDict()
        self.lora_B = nn.ModuleDict()
        self.scaling: Dict[str, float] = {}
        self.r: Dict[str, int] = {}
        self.lora_alpha: Dict[str, int] = {}
        self._active_adapter = []
        self._disable_adapters = False

    def update_layer(self, name: str, r: int, alpha: int, init: str = "gaussian"):
        out_f, in_f = self.base_layer.weight.shape
        dev = self.base_layer.weight.device
        a, b = ParamNode(), ParamNode()
        if _EMPTY_INIT:
            wa = torch.empty(r, in_f, device=dev)
        elif init == "gaussian":
            wa = torch.randn(r, in_f, device=dev) / r
        else:
            wa = (torch.rand(r, in_f, device=dev) * 2 - 1) / math.sqrt(in_f)
        a.register_parameter("weight", nn.Parameter(wa))
        b.register_parameter("weight", nn.Parameter(torch.empty(out_f, r, device=dev) if _EMPTY_INIT else torch.zeros(out_f, r, device=dev)))
        self.lora_A[name] = a
        self.lora_B[name] = b
        self.r[name], self.lora_alpha[name], self.scaling[name] = r, alpha, alpha / r

    @property
    def active_adapters(self) -> List[str]:
        a = self._active_adapter
        return [a] if isinstance(a, str) else list(a)


class UNetParams(ParamNode):
    """Holder with the diffusers UNet2DConditionModel surface MADM touches (add_adapter / set_adapter)."""

    def __init__(self, device=None, seed: Optional[int] = None, in_channels: int = 4):
        super().__init__()
        tree = build_tree(unet_spec(in_channels=in_channels), device=device, seed=seed)
        for k, m in tree._modules.items():
            self.add_module(k, m)
        self.adapter_names: List[str] = []
        self._struct_version = 0  # bumped whenever the parameter SET changes (add_adapter): name / tensor lists are cached per version

    def lora_layers(self) -> Iterable[Tuple[str, LoraLinearParams]]:
        for n, m in self.named_modules():
            if isinstance(m, LoraLinearParams):
                yield n, m

    def add_adapter(self, adapter_config, adapter_name: str = "default"):
        r = int(getattr(adapter_config, "r"))
        alpha = int(getattr(adapter_config, "lora_alpha"))
        targets = tuple(getattr(adapter_config, "target_modules", LORA_TARGETS) or LORA_TARGETS)
        init = getattr(adapter_config, "init_lora_weights", "gaussian")
        names = [n for n, m in self.named_modules()
                 if any(n == t or n.endswith("." + t) for t in targets)
                 and (isinstance(m, LoraLinearParams) or "weight" in m._parameters)]
        for n in names:
            parent_name, _, leaf = n.rpartition(".")
            parent = self.get_submodule(parent_name) if parent_name else self
            cur = parent._modules[leaf]
            if not isinstance(cur, LoraLinearParams):
                cur = LoraLinearParams(cur)
                parent._modules[leaf] = cur
            cur.update_layer(adapter_name, r, alpha, init if isinstance(init, str) else "kaiming")
        if adapter_name not in self.adapter_names:
            self.adapter_names.append(adapter_name)
        object.__setattr__(self, "_struct_version", getattr(self, "_struct_version", 0) + 1)
        self.set_adapter(adapter_name)

    def set_adapter(self, adapter_name):
        names = [adapter_name] if isinstance(adapter_name, str) else list(adapter_name)
        for _, m in self.lora_layers():
            m._active_adapter = names

    def active_adapter(self) -> Optional[str]:
        """The single adapter the reference activates per forward (mtmadise.py:144-147); None = base weights."""
        for _, m in self.lora_layers():
            if m._disable_adapters:
                return None
            act = [a for a in m.active_adapters if a in m.lora_A]
            if len(act) > 1:
                raise NotImplementedError(
                    f"{len(act)} LoRA adapters active at once ({act}); MADM activates exactly one per forward "
                    "(MTMADISE.set_lora_adapter) and the folded-weight path follows that")
            return act[0] if act else None
        return None

    def scaling_of(self, adapter: str) -> float:
        for _, m in self.lora_layers():
            return m.scaling[adapter]
        return 0.0


class VAEParams(ParamNode):
    scaling_factor = 0.18215
    latent_channels = 4

    def __init__(self, device=None, seed: Optional[int] = None, with_decoder: bool = False):
        super().__init__()
        tree = build_tree(vae_spec() + (vae_decoder_spec() if with_decoder else []), device=device, seed=seed)
        for k, m in tree._modules.items():
            self.add_module(k, m)


class BottleneckParams(ParamNode):
    def __init__(self, cin: int, cout: int, mid: int, device=None):
        super().__init__()
        tree = build_tree(bottleneck_spec(cin, cout, mid), kind="d2", device=device)
        for k, m in tree._modules.items():
            self.add_module(k, m)
        self.in_channels, self.out_channels = cin, cout

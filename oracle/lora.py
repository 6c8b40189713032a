"""peft==0.10.0 LoRA ``Linear`` semantics as MADM uses them (reference
``modeling/meta_arch/mtmadise.py:115-147``).  TEST INFRASTRUCTURE; PARITY UNPINNED (peft absent).

``add_adapter`` wraps every ``to_q/to_k/to_v/to_out.0`` of the UNet (attn1 and attn2 of the 16
transformer blocks = 128 layers) so state-dict keys become ``...to_q.base_layer.weight``,
``...to_q.lora_A.<adapter>.weight [r,in]``, ``...to_q.lora_B.<adapter>.weight [out,r]``.
Forward (unmerged, dropout 0): ``y = base(x) + sum_{a in active} lora_B[a](lora_A[a](x)) * alpha/r``.
The reference writes a one-element list into ``module._active_adapter`` per forward
(``mtmadise.py:144-147``).
"""
from typing import Dict, List, Union

import torch
import torch.nn as nn

LORA_TARGETS = ("to_k", "to_q", "to_v", "to_out.0")


class LoraLinear(nn.Module):
    def __init__(self, base: nn.Linear):
        super().__init__()
        self.base_layer = base
        self.lora_A = nn.ModuleDict()
        self.lora_B = nn.ModuleDict()
        self.scaling: Dict[str, float] = {}
        self._active_adapter: Union[str, List[str]] = []
        self._disable_adapters = False

    @property
    def in_features(self):
        return self.base_layer.in_features

    @property
    def out_features(self):
        return self.base_layer.out_features

    def update_layer(self, name: str, r: int, alpha: int):
        self.lora_A[name] = nn.Linear(self.in_features, r, bias=False)
        self.lora_B[name] = nn.Linear(r, self.out_features, bias=False)
        self.scaling[name] = alpha / r
        # init_lora_weights="gaussian": A ~ N(0, 1/r), B = 0
        nn.init.normal_(self.lora_A[name].weight, std=1.0 / r)
        nn.init.zeros_(self.lora_B[name].weight)

    @property
    def active_adapters(self) -> List[str]:
        a = self._active_adapter
        return [a] if isinstance(a, str) else list(a)

    def forward(self, x):
        y = self.base_layer(x)
        if self._disable_adapters:
            return y
        for a in self.active_adapters:
            if a in self.lora_A:
                y = y + self.lora_B[a](self.lora_A[a](x)) * self.scaling[a]
        return y

    def merged_weight(self, adapter: str) -> torch.Tensor:
        """W' = W + (alpha/r) * B @ A  (what a weight-folding implementation must equal)."""
        return self.base_layer.weight + self.scaling[adapter] * (self.lora_B[adapter].weight @ self.lora_A[adapter].weight)


def _get_parent(root: nn.Module, dotted: str):
    parts = dotted.split(".")
    m = root
    for p in parts[:-1]:
        m = m[int(p)] if p.isdigit() else getattr(m, p)
    return m, parts[-1]


def add_adapter(unet: nn.Module, name: str, r: int, alpha: int):
    """unet.add_adapter(LoraConfig(r, lora_alpha, init 'gaussian', target_modules=LORA_TARGETS), name)."""
    targets = []
    for mod_name, mod in unet.named_modules():
        if any(mod_name == t or mod_name.endswith("." + t) for t in LORA_TARGETS):
            if isinstance(mod, (nn.Linear, LoraLinear)):
                targets.append(mod_name)
    for mod_name in targets:
        parent, leaf = _get_parent(unet, mod_name)
        cur = parent[int(leaf)] if leaf.isdigit() else getattr(parent, leaf)
        if not isinstance(cur, LoraLinear):
            cur = LoraLinear(cur)
            if leaf.isdigit():
                parent[int(leaf)] = cur
            else:
                setattr(parent, leaf, cur)
        cur.update_layer(name, r, alpha)
    return len(targets)


def set_adapter(unet: nn.Module, state):
    """MTMADISE.set_lora_adapter: write ``_active_adapter`` on every tuner layer (mtmadise.py:129-147)."""
    if isinstance(state, str):
        state = [state]
    for _, m in unet.named_modules():
        if isinstance(m, LoraLinear):
            m._active_adapter = state

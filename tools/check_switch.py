#!/usr/bin/env python
"""One forward of the seeded synthetic product model (no oracle) under whatever MADM_* debug switches are in the environment; writes the
feature maps to the given .pt file.  Used by tests/test_switches_gpu.py to check that every switch still computes the same thing."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from helpers import build_product_backbone, set_lora_adapter  # noqa: E402

out_path, variant = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "base")
dev = torch.device("cuda:0")
torch.manual_seed(1234)
bb = build_product_backbone(dev, variant=variant)
g = torch.Generator(device=dev).manual_seed(5)
with torch.no_grad():
    for n, p in sorted(bb.named_parameters()):  # non-trivial norm affines / LoRA B like the parity fixtures
        if "norm" in n and n.endswith("weight"):
            p.add_(0.1 * torch.randn(p.shape, device=dev, generator=g))
        elif "lora_B" in n:
            p.copy_(0.02 * torch.randn(p.shape, device=dev, generator=g))
set_lora_adapter(bb.feature_extractor.ldm_extractor.unet, "Depth")
img = torch.rand(2, 3, 512, 512, device=dev, generator=g)
with torch.no_grad():
    feats = bb(img, input_modal="others")["output_features"]
torch.save({k: v.float().cpu() for k, v in feats.items()}, out_path)
print("ok", {k: tuple(v.shape) for k, v in feats.items()})

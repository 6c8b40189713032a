#!/usr/bin/env python
"""Gradients of the seeded synthetic product model (no oracle) under whatever MADM_* switches are in the environment -> .pt file;
`check_grads.py cmp a.pt b.pt` prints per-family cosine / norm ratio of two such files (A/B of kernel paths in the backward pass)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

if sys.argv[1] == "cmp":
    a, b = torch.load(sys.argv[2]), torch.load(sys.argv[3])
    fam = {}
    for n in a:
        f = "feat" if n.startswith("feat:") else ("lora_A" if "lora_A" in n else "lora_B" if "lora_B" in n else "proj" if n.startswith("feature_projections.") else "cond")
        x, y = fam.setdefault(f, ([], []))
        x.append(a[n].flatten().double()); y.append(b[n].flatten().double())
    for f, (x, y) in sorted(fam.items()):
        x, y = torch.cat(x), torch.cat(y)
        print(f"{f:8s} cosine {torch.nn.functional.cosine_similarity(x, y, dim=0).item():.7f}  |a|/|b| {(x.norm() / y.norm()).item():.5f}  rel diff {((x - y).norm() / y.norm()).item():.2e}")
    sys.exit(0)

from helpers import build_product_backbone, set_lora_adapter  # noqa: E402
from test_train_gpu import make_trainable, product_loss  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(1234)
bb = build_product_backbone(dev, compute_dtype="fp16")
g = torch.Generator(device=dev).manual_seed(5)
with torch.no_grad():
    for n, p in sorted(bb.named_parameters()):
        if "norm" in n and n.endswith("weight"):
            p.add_(0.1 * torch.randn(p.shape, device=dev, generator=g))
        elif "lora_B" in n:
            p.copy_(0.02 * torch.randn(p.shape, device=dev, generator=g))
make_trainable(bb)
set_lora_adapter(bb.feature_extractor.ldm_extractor.unet, "Depth")
img = torch.rand(2, 3, 512, 512, device=dev, generator=g)
out = bb(img, input_modal="others")["output_features"]
product_loss(out).backward()
res = {"feat:" + k: v.detach().float().cpu() for k, v in out.items()}
res.update({n: p.grad.float().cpu() for n, p in bb.named_parameters() if p.grad is not None})
torch.save(res, sys.argv[1])
print("ok", len(res))

#!/usr/bin/env python
"""Per-stage (VAE / UNet / projections) and per-kernel-family device time of one madm_extract step (CUDA events recorded by the
engine around every launch)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from helpers import build_product_backbone, set_lora_adapter  # noqa: E402
from madm_b200 import _lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
bb = build_product_backbone(dev)
ldm = bb.feature_extractor.ldm_extractor
set_lora_adapter(ldm.unet, "Depth")
img = torch.rand(B, 3, 512, 512, device=dev)
eng = ldm.engine()
eng.graph_max_batch = 0
with torch.no_grad():
    for _ in range(3):
        bb._extract(img, "others", False, None)
    batched = dict(img=img)
    bb.feature_extractor.conditioning(batched, "others", False, None)
    eng.set_profiling(True)
    tot = {}
    for name, st in (("vae", _lib.STAGE_VAE), ("unet", _lib.STAGE_UNET), ("proj", _lib.STAGE_PROJ)):
        ldm.run(batched, "others", stages=st, extra=bb._projection_tensors())
        p = eng.profile()
        tot[name] = p
        print(name, "total %.3f ms |" % sum(v["ms"] for v in p.values()),
              "  ".join(f"{k} {v['ms']:.3f}ms/{int(v['launches'])} ({v['flops'] / max(v['ms'], 1e-9) / 1e9:.0f} TF/s)" if v["flops"] else
                       f"{k} {v['ms']:.3f}ms/{int(v['launches'])} ({v['bytes'] / max(v['ms'], 1e-9) / 1e6:.0f} GB/s)" for k, v in p.items()))

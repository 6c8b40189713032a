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
VARIANT = sys.argv[2] if len(sys.argv) > 2 else "base"  # "s0": the vae_decoder_loss variant
dev = torch.device("cuda:0")
bb = build_product_backbone(dev, variant=VARIANT)
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
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        bb._extract(img, "others", False, None)
    e1.record()
    torch.cuda.synchronize()
    print(f"{VARIANT} B={B}: {e0.elapsed_time(e1) / 5:.3f} ms per step (eager launches), {B * 5000 / e0.elapsed_time(e1):.1f} img/s,",
          f"workspace {eng._ws.numel() / 2**30:.2f} GiB, packed {eng._packed.numel() / 2**30:.2f} GiB")
    stages = [("vae", _lib.STAGE_VAE), ("unet", _lib.STAGE_UNET)] + ([("dec", _lib.STAGE_DEC)] if VARIANT == "s0" else []) + [("proj", _lib.STAGE_PROJ)]
    for name, st in stages:
        eng.extract(img if st == _lib.STAGE_VAE else None, batched["cond_inputs"].expand(B, -1, -1), batched["cond_emb"][:, 0].expand(B, -1),
                    torch.zeros(B, dtype=torch.int64, device=dev), ldm.shared_noise, stages=st, B=B)
        p = eng.profile()
        tot[name] = p
        print(name, "total %.3f ms |" % sum(v["ms"] for v in p.values()),
              "  ".join(f"{k} {v['ms']:.3f}ms/{int(v['launches'])} ({v['flops'] / max(v['ms'], 1e-9) / 1e9:.0f} TF/s)" if v["flops"] else
                       f"{k} {v['ms']:.3f}ms/{int(v['launches'])} ({v['bytes'] / max(v['ms'], 1e-9) / 1e6:.0f} GB/s)" for k, v in p.items()))

#!/usr/bin/env python
"""One profiled step of the hot path for ncu: builds the synthetic model, runs warm-up steps, then brackets exactly
`--steps` madm_extract calls with cudaProfilerStart/Stop (run ncu with --profile-from-start off)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from helpers import build_product_backbone, set_lora_adapter  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--dtype", default="fp16")
ap.add_argument("--variant", default="base", choices=["base", "s0"])
args = ap.parse_args()
dev = torch.device("cuda:0")
bb = build_product_backbone(dev, compute_dtype=args.dtype, variant=args.variant)
set_lora_adapter(bb.feature_extractor.ldm_extractor.unet, "Depth")
bb.feature_extractor.ldm_extractor.engine().graph_max_batch = 0  # profile the stream launches, not a graph replay
img = torch.rand(args.batch, 3, 512, 512, device=dev)
with torch.no_grad():
    for _ in range(2):
        bb._extract(img, "others", False, None)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    for _ in range(args.steps):
        bb._extract(img, "others", False, None)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("profiled", args.steps, "step(s), batch", args.batch)

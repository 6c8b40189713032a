#!/usr/bin/env python
"""One LoRA training pass of the backbone (forward under grad + backward) for profiling.
  python tools/profile_train.py cpu    : cProfile of 3 passes (host-side cost of the autograd.Function / planner calls)
  ncu --profile-from-start off ... python tools/profile_train.py ncu : exactly one pass between cudaProfilerStart/Stop
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from helpers import build_product_backbone, set_lora_adapter  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "cpu"
dtype = sys.argv[2] if len(sys.argv) > 2 else "fp16"
dev = torch.device("cuda:0")
bb = build_product_backbone(dev, compute_dtype=dtype)
for n, p in bb.named_parameters():
    p.requires_grad_(("lora_" in n) or n.startswith("feature_projections.") or n.startswith("feature_extractor.clip_project_"))
set_lora_adapter(bb.feature_extractor.ldm_extractor.unet, "Depth")
img = torch.rand(2, 3, 512, 512, device=dev)
R = [torch.randn(2, 512, s, s, device=dev) / 1e3 for s in (128, 64, 32, 16)]


def one_pass():
    out = bb(img, input_modal="others")["output_features"]
    loss = sum((f * r).sum() for f, r in zip(out.values(), R))
    loss.backward()
    for p in bb.parameters():
        p.grad = None


for _ in range(3):
    one_pass()
torch.cuda.synchronize()
if mode == "cpu":
    import cProfile
    import pstats
    import time
    t0 = time.perf_counter()
    for _ in range(3):
        one_pass()
    torch.cuda.synchronize()
    print(f"wall per pass: {(time.perf_counter() - t0) / 3 * 1e3:.1f} ms")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(3):
        one_pass()
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
else:
    torch.cuda.cudart().cudaProfilerStart()
    one_pass()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("profiled one training pass")

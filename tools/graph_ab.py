#!/usr/bin/env python
"""A/B at B = 8: stream launches (~555 per step) vs one CUDA-graph replay of the same step."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from helpers import build_product_backbone, set_lora_adapter  # noqa: E402

dev = torch.device("cuda:0")
bb = build_product_backbone(dev)
ldm = bb.feature_extractor.ldm_extractor
set_lora_adapter(ldm.unet, "Depth")
img = torch.rand(8, 3, 512, 512, device=dev)
eng = ldm.engine()


def timed(n=10):
    for _ in range(3):
        bb._extract(img, "others", False, None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        bb._extract(img, "others", False, None)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    eng.graph_max_batch = 0
    a = timed()
    eng.graph_max_batch = 8
    b = timed()
print(f"B=8 stream launches {a:.3f} ms/step ({8e3 / a:.1f} img/s)   graph replay {b:.3f} ms/step ({8e3 / b:.1f} img/s)")

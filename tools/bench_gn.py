#!/usr/bin/env python
"""Micro-benchmark of the GroupNorm passes through the C ABI (inputs larger than L2; CUDA events on the launch stream).
Prints time and algorithmic GB/s of (statistics + apply) and of (column-statistics reduce + apply) per shape."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from madm_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
shapes = [(8, 512 * 512, 128), (8, 256 * 256, 256), (8, 128 * 128, 512), (8, 64 * 64, 512), (8, 64 * 64, 320), (8, 64 * 64, 960),
          (8, 32 * 32, 640), (8, 16 * 16, 1280), (8, 8 * 8, 2560)]


def timed(fn, n=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for B, HW, C in shapes:
    for in16 in (False, True):
        x = torch.randn(B, HW, C, device=dev, dtype=torch.float16 if in16 else torch.float32)
        g = torch.randn(C, device=dev)
        bt = torch.randn(C, device=dev)
        y = torch.empty(B, HW, C, device=dev, dtype=torch.float16)
        nb = HW // 32
        cs = torch.randn(B * nb * C * 2, device=dev)
        t_full = timed(lambda: ops.groupnorm(x, None, B, HW, g, bt, 1e-5, 1, y))
        t_cs = timed(lambda: ops.groupnorm_from_colstats(x, B, HW, cs, 32, g, bt, 1e-5, 1, y))
        el = B * HW * C
        ib = 2 if in16 else 4
        print(f"B={B} HW={HW:7d} C={C:5d} in16={int(in16)}  stats+apply {t_full * 1e3:8.1f} us ({el * (2 * ib + 2) / t_full / 1e6:7.0f} GB/s)   "
              f"colstats+apply {t_cs * 1e3:8.1f} us ({el * (ib + 2) / t_cs / 1e6:7.0f} GB/s)")

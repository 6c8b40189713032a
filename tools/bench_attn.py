#!/usr/bin/env python
"""Micro-benchmark of the attention kernel through the C ABI on the UNet's shapes (fused-QKV layout, B=8, 8 heads)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from madm_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, heads = 8, 8
only = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for d, n, nk in ((40, 4096, 4096), (40, 4096, 77), (80, 1024, 1024), (160, 256, 256)):
    if only and d != only:
        continue
    C = heads * d
    q = torch.randn(B * n, 3 * C, device=dev, dtype=torch.float16)
    if nk == n:
        k, v, ldk, kvbs = q[:, C:], q[:, 2 * C:], 3 * C, n * 3 * C
    else:
        kv = torch.randn(B * nk, 2 * C, device=dev, dtype=torch.float16)
        k, v, ldk, kvbs = kv, kv[:, C:], 2 * C, nk * 2 * C
    o = torch.empty(B * n, C, device=dev, dtype=torch.float16)
    run = lambda: ops.attention(q, 3 * C, k, ldk, v, ldk, o, C, B, heads, d, n, nk, n * 3 * C, kvbs, n * C, d ** -0.5)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 4.0 * B * heads * n * nk * d
    print(f"d={d:3d} n={n:5d} keys={nk:5d}: {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s (algorithmic, unpadded d)")

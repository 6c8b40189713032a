#!/usr/bin/env python
"""Run-to-run bit-exactness of the GEMM and attention kernels on layer shapes of the path (no atomics -> must be identical)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from madm_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
DT = torch.float16
bad = 0
for name, B, H, W, Ci, Co, taps in [("conv 256^2 128->128", 8, 256, 256, 128, 128, 9), ("conv 128^2 256->256", 8, 128, 128, 256, 256, 9),
                                    ("conv 64^2 320->320", 8, 64, 64, 320, 320, 9), ("conv 16^2 1280->1280", 8, 16, 16, 1280, 1280, 9),
                                    ("lin 64^2 320->960", 8, 64, 64, 320, 960, 1), ("lin 64^2 1280->320", 8, 64, 64, 1280, 320, 1)]:
    M = B * H * W
    x = (torch.randn(B, H, W, Ci, device=dev) * 0.5).to(DT)
    w = (torch.randn(Co, taps * Ci, device=dev) * 0.02).to(DT)
    seg = ops.make_seg(x, B, H, W, Ci, taps=ops.taps_3x3() if taps == 9 else None)
    outs = []
    for r in range(6):
        o = torch.empty(M, Co, device=dev)
        ops.gemm([seg], M, Co, w, out_f32=o, ldo32=Co)
        outs.append(o)
    torch.cuda.synchronize()
    nd = sum(int(not torch.equal(outs[0], o)) for o in outs[1:])
    ref = (x.float().reshape(M, Ci) @ w.float().t()) if taps == 1 else None
    err = float((outs[0] - ref).abs().max()) if ref is not None else float("nan")
    print(f"gemm {name:24s}: {nd} of 5 repeats differ   max|err| vs torch {err:.3e}")
    bad += nd
for d, n, nk in ((40, 4096, 4096), (40, 4096, 77), (80, 1024, 1024), (160, 256, 256)):
    heads, C = 8, 8 * d
    q = torch.randn(8 * n, 3 * C, device=dev, dtype=DT)
    if nk == n:
        k, v, ldk, kvbs = q[:, C:], q[:, 2 * C:], 3 * C, n * 3 * C
    else:
        kv = torch.randn(8 * nk, 2 * C, device=dev, dtype=DT)
        k, v, ldk, kvbs = kv, kv[:, C:], 2 * C, nk * 2 * C
    outs = []
    for r in range(6):
        o = torch.empty(8 * n, C, device=dev, dtype=DT)
        ops.attention(q, 3 * C, k, ldk, v, ldk, o, C, 8, heads, d, n, nk, n * 3 * C, kvbs, n * C, d ** -0.5)
        outs.append(o)
    torch.cuda.synchronize()
    nd = sum(int(not torch.equal(outs[0], o)) for o in outs[1:])
    print(f"attention d={d} n={n} keys={nk}: {nd} of 5 repeats differ")
    bad += nd
sys.exit(1 if bad else 0)

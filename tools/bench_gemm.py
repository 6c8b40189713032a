#!/usr/bin/env python
"""Micro-benchmark of the implicit-GEMM kernel through the C ABI on layer shapes of the path (3x3 convs and 1x1 / linear)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from madm_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
DT = torch.float16
# (name, B, H, W, Cin, Cout, taps, out16, out32, residual)
shapes = [
    ("vae conv 512^2 128->128", 8, 512, 512, 128, 128, 9, True, False, False),
    ("vae conv 512^2 128->128 +res f32", 8, 512, 512, 128, 128, 9, True, True, True),
    ("vae conv 256^2 256->256", 8, 256, 256, 256, 256, 9, True, False, False),
    ("vae conv 128^2 512->512", 8, 128, 128, 512, 512, 9, True, False, False),
    ("unet conv 64^2 320->320", 8, 64, 64, 320, 320, 9, True, False, False),
    ("unet conv 32^2 640->640", 8, 32, 32, 640, 640, 9, True, False, False),
    ("unet conv 16^2 1280->1280", 8, 16, 16, 1280, 1280, 9, True, False, False),
    ("unet lin 64^2 320->320 +res f32", 8, 64, 64, 320, 320, 1, False, True, True),
    ("unet lin 64^2 320->960 (qkv)", 8, 64, 64, 320, 960, 1, True, False, False),
    ("unet lin 64^2 1280->320 +res", 8, 64, 64, 1280, 320, 1, True, False, True),
]
sel = sys.argv[1] if len(sys.argv) > 1 else ""
for name, B, H, W, Ci, Co, taps, o16, o32, res in shapes:
    if sel and sel not in name:
        continue
    M = B * H * W
    x = (torch.randn(B, H, W, Ci, device=dev) * 0.5).to(DT)
    w = (torch.randn(Co, taps * Ci, device=dev) * 0.02).to(DT)
    bias = torch.randn(Co, device=dev)
    out16 = torch.empty(M, Co, device=dev, dtype=DT) if o16 else None
    out32 = torch.empty(M, Co, device=dev) if (o32 or not o16) else None
    resid = torch.randn(M, Co, device=dev) if res else None
    seg = ops.make_seg(x, B, H, W, Ci, taps=ops.taps_3x3() if taps == 9 else None)
    kw = dict(bias=bias)
    if out32 is not None:
        kw.update(out_f32=out32, ldo32=Co)
    if out16 is not None:
        kw.update(out_bf16=out16, ldo16=Co)
    if resid is not None:
        kw.update(residual=resid, ldr=Co)
    run = lambda: ops.gemm([seg], M, Co, w, **kw)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{name:36s} M={M:8d} N={Co:5d} K={taps * Ci:6d}: {ms * 1e3:8.1f} us  {2.0 * M * Co * taps * Ci / ms / 1e9:7.1f} TFLOP/s", flush=True)

#!/usr/bin/env python
"""Output-bound GEMMs (short K): the epilogue's store path is what is measured.  VAE conv_in shape (M = 2.1M, K = 64, N = 128), the UNet
short-K linears, each as 16-bit out / fp32 out / fp32 residual in place, with and without fused column statistics.
With a -DGEMM_INSTR build (MADM_B200_LIB=...) the kernel prints its epilogue cycle counters."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from madm_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
DT = torch.float16
only = sys.argv[1] if len(sys.argv) > 1 else ""
print("MADM_GEMM_TMA_EPI =", os.environ.get("MADM_GEMM_TMA_EPI", "(default: on)"), flush=True)
for name, M, K, N, mode, stats in [("conv_in 2.1Mx128x64 h16", 2097152, 64, 128, "h16", False), ("conv_in 2.1Mx128x64 h16+stats", 2097152, 64, 128, "h16", True),
                                   ("2.1Mx128x128 f32+stats", 2097152, 128, 128, "f32", True), ("32768x320x320 res", 32768, 320, 320, "res", True),
                                   ("32768x320x320 h16", 32768, 320, 320, "h16", False), ("8192x640x640 res", 8192, 640, 640, "res", True),
                                   ("32768x960x320 h16 (qkv)", 32768, 320, 960, "h16", False), ("32768x320x320 res plain", 32768, 320, 320, "res", False),
                                   ("8192x640x640 res plain", 8192, 640, 640, "res", False), ("2048x1280x1280 res plain", 2048, 1280, 1280, "res", False),
                                   ("32768x320x320 f32", 32768, 320, 320, "f32", False), ("8192x1920x640 h16 (qkv)", 8192, 640, 1920, "h16", False),
                                   ("2.1Mx128x128 f32 plain", 2097152, 128, 128, "f32", False), ("32768x320x1280 res->h16 (ff out)", 32768, 1280, 320, "resh16", False),
                                   ("8192x640x2560 res->h16 (ff out)", 8192, 2560, 640, "resh16", False), ("2048x1280x5120 res->h16 (ff out)", 2048, 5120, 1280, "resh16", False)]:
    if only and only not in name:
        continue
    x = torch.randn(M, K, device=dev).to(DT)
    w = (torch.randn(N, K, device=dev) * 0.05).to(DT)
    o16 = torch.empty(M, N, device=dev, dtype=DT) if mode in ("h16", "resh16") else None
    o32 = torch.randn(M, N, device=dev) if mode != "h16" else None
    cs = torch.empty((M + 31) // 32, N, 2, device=dev) if stats else None
    seg = ops.make_seg(x, 1, 1, M, K)
    kw = dict(out_bf16=o16, ldo16=N) if mode in ("h16", "resh16") else dict(out_f32=o32, ldo32=N)
    if mode in ("res", "resh16"):
        kw.update(residual=o32, ldr=N)
    run = lambda: ops.gemm([seg], M, N, w, colstats=cs, stat_rows=32 if stats else 0, **kw)
    run(); torch.cuda.synchronize()
    n = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        run()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / n * 1e3
    by = M * K * 2 + M * N * (2 if mode == "h16" else 4) * (2 if mode == "res" else 1)
    print(f"{name:34s} {t:8.1f} us/launch  {by / t / 1e6:6.2f} TB/s algorithmic  {2.0 * M * N * K / t / 1e6:7.1f} TFLOP/s", flush=True)

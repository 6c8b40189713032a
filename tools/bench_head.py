#!/usr/bin/env python
"""DAFormer head stage at B = 8 through the public module: device time per call and per kernel family."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from madm_b200.head import DAFormerHead  # noqa: E402
from test_head_gpu import HEAD_KW  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
VARIANT = sys.argv[2] if len(sys.argv) > 2 else "base"  # "s0": in_keys[0]='s0', in_channels[0]=128, fused on the 512^2 grid
if VARIANT == "s0":
    head = DAFormerHead(**dict(HEAD_KW, in_channels=[128, 512, 512, 512], in_keys=["s0", "s3", "s4", "s5"]), device=dev).eval()
    feats = {k: F.relu(torch.randn(B, c, s, s, device=dev)) for k, c, s in zip(("s0", "s3", "s4", "s5"), (128, 512, 512, 512), (512, 64, 32, 16))}
else:
    head = DAFormerHead(**HEAD_KW, device=dev).eval()
    feats = {k: F.relu(torch.randn(B, 512, s, s, device=dev)) for k, s in zip(("s2", "s3", "s4", "s5"), (128, 64, 32, 16))}
with torch.no_grad():
    for _ in range(3):
        head({"output_features": feats})
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        head({"output_features": feats})
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    eng = head.engine()
    eng.set_profiling(True)
    head({"output_features": feats})
    prof = eng.profile()
    eng.set_profiling(False)
gfi = 1890.0 if VARIANT == "s0" else 118.0  # ~118 GFLOP per image on the 128^2 grid, ~1.9 TFLOP on the 512^2 grid
gf = gfi * B
print(f"DAFormer head ({VARIANT}) B={B}: {ms:.3f} ms/call ({B / ms * 1e3:.0f} img/s, ~{gf / ms:.0f} TFLOP/s on ~{gfi:.0f} GFLOP/img)")
for k, v in prof.items():
    if v["launches"]:
        extra = f"{v['flops'] / v['ms'] / 1e9:.0f} TFLOP/s" if v["flops"] else f"{v['bytes'] / v['ms'] / 1e6:.0f} GB/s"
        print(f"   {k:16s} {v['ms']:.3f} ms / {int(v['launches'])} launches  ({extra})")

#!/usr/bin/env python
"""CTA-pair (cta_group::2) GEMM against the single-CTA kernel on the same inputs: must be bit-identical; prints timings."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from madm_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
DT = torch.float16
cases = [  # name, B, H, W, Cin, Cout, taps, bn, mt
    ("lin M=640 (odd tiles) 320->320", 1, 1, 640, 320, 320, 1, 160, 0),
    ("lin 64^2 320->960", 8, 64, 64, 320, 960, 1, 0, 0),
    ("conv 64^2 320->320", 8, 64, 64, 320, 320, 9, 0, 0),
    ("conv 32^2 640->640", 8, 32, 32, 640, 640, 9, 0, 0),
    ("conv 128^2 256->256", 8, 128, 128, 256, 256, 9, 0, 0),
    ("conv 128^2 512->512", 8, 128, 128, 512, 512, 9, 0, 0),
    ("conv 256^2 128->128 (mt=2)", 8, 256, 256, 128, 128, 9, 128, 2),
    ("conv 512^2 128->128 (mt=2)", 8, 512, 512, 128, 128, 9, 128, 2),
]
sel = sys.argv[1] if len(sys.argv) > 1 else ""
bad = 0
for name, B, H, W, Ci, Co, taps, bn, mt in cases:
    if sel and sel not in name:
        continue
    M = B * H * W
    x = (torch.randn(B, H, W, Ci, device=dev) * 0.5).to(DT)
    w = (torch.randn(Co, taps * Ci, device=dev) * 0.02).to(DT)
    bias = torch.randn(Co, device=dev)
    seg = ops.make_seg(x, B, H, W, Ci, taps=ops.taps_3x3() if taps == 9 else None) if H > 1 else ops.make_seg(x.reshape(M, Ci), 1, 1, M, Ci)
    res = {}
    for pair in (-1, 1):
        o32 = torch.empty(M, Co, device=dev)
        o16 = torch.empty(M, Co, device=dev, dtype=DT)
        run = lambda: ops.gemm([seg], M, Co, w, bias=bias, out_f32=o32, ldo32=Co, out_bf16=o16, ldo16=Co, bn=bn, mt=mt, pair=pair)
        run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            run()
        e1.record()
        torch.cuda.synchronize()
        res[pair] = (o32.clone(), o16.clone(), e0.elapsed_time(e1) / 5)
    same = torch.equal(res[-1][0], res[1][0]) and torch.equal(res[-1][1], res[1][1])
    md = float((res[-1][0] - res[1][0]).abs().max())
    fl = 2.0 * M * Co * taps * Ci
    print(f"{name:32s} identical={same} maxdiff={md:.2e}  single {res[-1][2] * 1e3:7.1f} us ({fl / res[-1][2] / 1e9:6.0f} TF/s)   pair {res[1][2] * 1e3:7.1f} us ({fl / res[1][2] / 1e9:6.0f} TF/s)", flush=True)
    bad += int(not same)
sys.exit(1 if bad else 0)

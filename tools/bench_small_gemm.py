#!/usr/bin/env python
"""Fixed cost of a GEMM launch: back-to-back tiny / short-K launches (stream and CUDA-graph replay), single CTA vs CTA pairs."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from madm_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
DT = torch.float16
for name, M, K, N, pair in [("tiny 128x128x64", 128, 64, 128, -1), ("one wave 18944x320x320", 18944, 320, 320, -1),
                            ("32768x320x320 single", 32768, 320, 320, -1), ("32768x320x320 pair", 32768, 320, 320, 1),
                            ("32768x320x1280 pair", 32768, 1280, 320, 1), ("8192x640x640 pair", 8192, 640, 640, 1), ("2048x1280x1280", 2048, 1280, 1280, 0)]:
    x = torch.randn(M, K, device=dev).to(DT)
    w = (torch.randn(N, K, device=dev) * 0.05).to(DT)
    res = torch.randn(M, N, device=dev)
    out = torch.empty(M, N, device=dev)
    seg = ops.make_seg(x, 1, 1, M, K)
    run = lambda: ops.gemm([seg], M, N, w, residual=res, ldr=N, out_f32=out, ldo32=N, pair=pair)
    run(); torch.cuda.synchronize()
    n = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        run()
    e1.record(); torch.cuda.synchronize()
    t_stream = e0.elapsed_time(e1) / n * 1e3
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        run()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        for _ in range(n):
            run()
    g.replay(); torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    t_graph = e0.elapsed_time(e1) / n * 1e3
    print(f"{name:28s} stream {t_stream:7.1f} us/launch   graph {t_graph:7.1f} us/launch   ({2.0 * M * N * K / t_graph / 1e6:7.1f} TFLOP/s in graph)")

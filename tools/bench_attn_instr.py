#!/usr/bin/env python
"""One launch of the d = 40 self-attention shape with a -DFA_INSTR build (MADM_B200_LIB=...): the kernel prints the softmax warps' phase cycle counters."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from madm_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, heads = 8, 8
d, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (40, 4096)
C = heads * d
q = torch.randn(B * n, 3 * C, device=dev, dtype=torch.float16)
o = torch.empty(B * n, C, device=dev, dtype=torch.float16)
ops.attention(q, 3 * C, q[:, C:], 3 * C, q[:, 2 * C:], 3 * C, o, C, B, heads, d, n, n, n * 3 * C, n * 3 * C, n * C, d ** -0.5)
torch.cuda.synchronize()

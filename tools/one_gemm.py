"""Run one GEMM shape a few times (target for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mdhs_b200  # noqa
from mdhs_b200 import ops

M, N, K = [int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (401408, 256, 64))]
bn = int(sys.argv[4]) if len(sys.argv) > 4 else 0
a = torch.randn(M, K, device="cuda").bfloat16()
b = torch.randn(N, K, device="cuda").bfloat16()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.gemm(a, b, out=out, bn_hint=bn)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))

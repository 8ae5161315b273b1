"""Depthwise 7x7 kernels at the ConvNeXt-Tiny stage shapes (B = 128): timing, and a target for ncu.
  python tools/one_dwconv.py [stage 1..4]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mdhs_b200  # noqa
from mdhs_b200 import ops

SHAPES = {1: (56, 96), 2: (28, 192), 3: (14, 384), 4: (7, 768)}
B = 128
which = [int(sys.argv[1])] if len(sys.argv) > 1 else [1, 2, 3, 4]
for s in which:
    HW, C = SHAPES[s]
    x = torch.randn(B * HW * HW, C, device="cuda").bfloat16()
    dy = torch.randn(B * HW * HW, C, device="cuda").bfloat16()
    w = torch.randn(C, 1, 7, 7, device="cuda")
    bias = torch.randn(C, device="cuda")
    dw, db = torch.zeros_like(w), torch.zeros_like(bias)

    def t(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    mb = x.numel() * 2 / 1e6
    print(f"stage {s}: {HW}x{HW}x{C} ({mb:.0f} MB)  fwd {t(lambda: ops.dwconv7(x, w, bias, B, HW, HW)):.0f} us  "
          f"dgrad {t(lambda: ops.dwconv7(dy, w, None, B, HW, HW, flip=True)):.0f} us  "
          f"wgrad {t(lambda: ops.dwconv7_wgrad(x, dy, dw, db, B, HW, HW)):.0f} us", flush=True)

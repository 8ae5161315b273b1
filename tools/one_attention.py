"""Attention kernels at the BERT-base shapes of the bench configs (B = 128, H = 12, D = 64): timing, and a target for ncu.
  python tools/one_attention.py [S ...]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mdhs_b200  # noqa
from mdhs_b200 import ops

B, H, D = 128, 12, 64
for S in ([int(a) for a in sys.argv[1:]] or [64, 128, 256, 512]):
    qkv = torch.randn(B * S, 3 * H * D, device="cuda").bfloat16()
    q, k, v = qkv[:, :H * D], qkv[:, H * D:2 * H * D], qkv[:, 2 * H * D:]
    mask = (torch.arange(S, device="cuda")[None, :] < torch.randint(S // 2, S + 1, (B, 1), device="cuda")).to(torch.uint8)
    do = torch.randn(B * S, H * D, device="cuda").bfloat16()
    dqkv = torch.empty_like(qkv)
    scale = 1.0 / math.sqrt(D)
    o, lse = ops.attention_fwd(q, k, v, B, H, S, S, D, scale, key_mask=mask, drop_p=0.1, seed=3)

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
    tf = t(lambda: ops.attention_fwd(q, k, v, B, H, S, S, D, scale, key_mask=mask, drop_p=0.1, seed=3))
    tb = t(lambda: ops.attention_bwd(q, k, v, o, do, lse, B, H, S, S, D, scale, key_mask=mask, drop_p=0.1, seed=3,
                                     dq=dqkv[:, :H * D], dk=dqkv[:, H * D:2 * H * D], dv=dqkv[:, 2 * H * D:]))
    gf = 4.0 * B * H * S * S * D / 1e9
    print(f"S={S}: fwd {tf:.0f} us ({gf / tf * 1e3:.0f} TFLOP/s)  bwd {tb:.0f} us ({2.5 * gf / tb * 1e3:.0f} TFLOP/s of the 5-product count)",
          flush=True)

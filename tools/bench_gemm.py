"""Micro-benchmark of mdhs_gemm_bf16 on the shapes of the hot path (CUDA-event timed, L2 flushed)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mdhs_b200  # noqa
from mdhs_b200 import ops

SHAPES = [  # (name, M, N, K, a_mn, b_mn, out fp32 accumulate split)
    ("bert_qkv", 8192, 2304, 768, False, False, 0),
    ("bert_out", 8192, 768, 768, False, False, 0),
    ("bert_ffn1", 8192, 3072, 768, False, False, 0),
    ("bert_ffn2", 8192, 768, 3072, False, False, 0),
    ("bert_ffn1_dgrad", 8192, 768, 3072, False, True, 0),
    ("bert_ffn1_wgrad", 3072, 768, 8192, True, True, 4),
    ("conv1x1_64_256", 401408, 256, 64, False, False, 0),
    ("conv1x1_256_64", 401408, 64, 256, False, False, 0),
    ("conv3x3_l3_im2col", 25088, 256, 2304, False, False, 0),
    ("square_8192", 8192, 8192, 8192, False, False, 0),
]


def main():
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    res = []
    for name, M, N, K, a_mn, b_mn, split in SHAPES:
        a = torch.randn((K, M) if a_mn else (M, K), device="cuda").bfloat16()
        b = torch.randn((K, N) if b_mn else (N, K), device="cuda").bfloat16()
        if split:
            out = torch.zeros(M, N, device="cuda")
            fn = lambda: ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, out=out, accumulate=True, split_k=split)
        else:
            out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
            fn = lambda: ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, out=out)
        for _ in range(3):
            fn()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        A = a.float().t() if a_mn else a.float()
        # torch reference timing (cuBLAS) for context
        at = a.t().contiguous() if a_mn else a
        bt = b.t().contiguous() if b_mn else b
        for _ in range(3):
            torch.matmul(at, bt.t())
        tt = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(at, bt.t())
            e1.record()
            torch.cuda.synchronize()
            tt.append(e0.elapsed_time(e1))
        tcu = sorted(tt)[len(tt) // 2]
        fl = 2.0 * M * N * K
        by = 2.0 * (M * K + N * K) + (4.0 if split else 2.0) * M * N
        res.append({"name": name, "M": M, "N": N, "K": K, "ms": round(t, 4), "tflops": round(fl / t / 1e9, 1),
                    "gbs": round(by / t / 1e6, 1), "cublas_ms": round(tcu, 4), "cublas_tflops": round(fl / tcu / 1e9, 1)})
        print(json.dumps(res[-1]), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/bench_gemm.json", "w"), indent=1)


if __name__ == "__main__":
    main()

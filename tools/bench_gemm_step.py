"""Micro-benchmark of mdhs_gemm_bf16 with the SAME epilogue options the training step uses (BERT / ResNet shapes).

Each case is captured into a CUDA graph of REP back-to-back launches over ROT rotating buffer sets (so big operands
come from HBM like in the step) and timed with CUDA events.  Output: one JSON line per case.
  python tools/bench_gemm_step.py [filter-substring] [bn_hint]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mdhs_b200  # noqa
from mdhs_b200 import ops

REP, ROT = 12, 3
T, C, F = 8192, 768, 3072

# name, M, N, K, a_mn, b_mn, options
CASES = [
    ("plain_8192_768_768", T, C, C, 0, 0, {}),
    ("plain_8192_3072_768", T, F, C, 0, 0, {}),
    ("plain_8192_768_3072", T, C, F, 0, 0, {}),
    ("qkv_fwd", T, 3 * C, C, 0, 0, dict(bias=1)),
    ("out_fwd", T, C, C, 0, 0, dict(bias=1, residual=1, drop=0.1)),
    ("out_fwd_nodrop", T, C, C, 0, 0, dict(bias=1, residual=1)),
    ("ffn1_fwd", T, F, C, 0, 0, dict(bias=1, act=ops.ACT_GELU, aux_out=1)),
    ("ffn1_fwd_noaux", T, F, C, 0, 0, dict(bias=1, act=ops.ACT_GELU)),
    ("ffn2_fwd", T, C, F, 0, 0, dict(bias=1, residual=1, drop=0.1)),
    ("ffn1_fwd_deriv", T, F, C, 0, 0, dict(bias=1, act=ops.ACT_GELU_DERIV, aux_out=1)),
    ("ffn2_dgrad", T, F, C, 0, 1, dict(dact=ops.ACT_GELU, aux_in=1)),
    ("ffn2_dgrad_mul", T, F, C, 0, 1, dict(dact=ops.ACT_MUL, aux_in=1)),
    ("plain_8192_3072_768_bn256", T, F, C, 0, 0, dict(bn=256)),
    ("plain_8192_768_3072_bn256", T, C, F, 0, 0, dict(bn=256)),
    ("qkv_fwd_bn128", T, 3 * C, C, 0, 0, dict(bias=1, bn=128)),
    ("ffn2_dgrad_plain", T, F, C, 0, 1, {}),
    ("ffn1_dgrad", T, C, F, 0, 1, dict(residual=1)),
    ("out_dgrad", T, C, C, 0, 1, {}),
    ("qkv_dgrad", T, C, 3 * C, 0, 1, dict(residual=1)),
    ("ffn2_wgrad", C, F, T, 1, 1, dict(acc=1)),
    ("ffn1_wgrad", F, C, T, 1, 1, dict(acc=1)),
    ("qkv_wgrad", 3 * C, C, T, 1, 1, dict(acc=1)),
    ("out_wgrad", C, C, T, 1, 1, dict(acc=1)),
    ("l1_1x1_64_256", 401408, 256, 64, 0, 0, {}),
    ("l1_1x1_256_64", 401408, 64, 256, 0, 0, {}),
    ("l1_dgrad_256_64", 401408, 64, 256, 0, 1, {}),
    ("l1_dgrad_64_256", 401408, 256, 64, 0, 1, dict(residual=1)),
    ("l2_1x1_128_512", 100352, 512, 128, 0, 0, {}),
    ("l3_1x1_256_1024", 25088, 1024, 256, 0, 0, {}),
    ("l3_1x1_1024_256", 25088, 256, 1024, 0, 0, {}),
    ("l4_1x1_2048_512", 6272, 512, 2048, 0, 0, {}),
    ("l3_wgrad_1024_256", 1024, 256, 25088, 1, 1, dict(acc=1)),
    ("l1_wgrad_256_64", 256, 64, 401408, 1, 1, dict(acc=1)),
    # implicit-GEMM 3x3 convolutions (TMA im2col operand): conv = (B, H, W, C_in); fprop / stride-1 dgrad share the shape
    ("l1_conv3x3_fprop", 401408, 64, 576, 0, 0, dict(conv=(128, 56, 56, 64), colsum=1)),
    ("l2_conv3x3_fprop", 100352, 128, 1152, 0, 0, dict(conv=(128, 28, 28, 128), colsum=1)),
    ("l3_conv3x3_fprop", 25088, 256, 2304, 0, 0, dict(conv=(128, 14, 14, 256), colsum=1)),
    ("l4_conv3x3_fprop", 6272, 512, 4608, 0, 0, dict(conv=(128, 7, 7, 512), colsum=1)),
    ("l1_conv3x3_dgrad_stat", 401408, 64, 576, 0, 0, dict(conv=(128, 56, 56, 64), stat=1)),
    ("l3_conv3x3_dgrad_stat", 25088, 256, 2304, 0, 0, dict(conv=(128, 14, 14, 256), stat=1)),
    ("l1_conv3x3_wgrad", 64, 576, 401408, 1, 1, dict(acc=1, wconv=(128, 56, 56, 64))),
    ("l3_conv3x3_wgrad", 256, 2304, 25088, 1, 1, dict(acc=1, wconv=(128, 14, 14, 256))),
    ("l1_1x1_64_256_colsum", 401408, 256, 64, 0, 0, dict(colsum=1)),
    ("l1_dgrad_256_64_stat", 401408, 64, 256, 0, 1, dict(stat=1)),
    ("l3_dgrad_1024_256_stat", 25088, 256, 1024, 0, 1, dict(stat=1)),
]


def main():
    filt = sys.argv[1] if len(sys.argv) > 1 else ""
    bn_hint = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    dev = "cuda"
    for name, M, N, K, a_mn, b_mn, o in CASES:
        if filt and filt not in name:
            continue
        sets = []
        for _ in range(ROT):
            if o.get("conv"):
                cb, ch, cw, cc = o["conv"]
                a = torch.randn((cb * ch * cw, cc), device=dev).mul_(0.5).bfloat16()
            else:
                a = torch.randn((K, M) if a_mn else (M, K), device=dev).mul_(0.5).bfloat16()
            if o.get("wconv"):
                cb, ch, cw, cc = o["wconv"]
                b = torch.randn((cb * ch * cw, cc), device=dev).mul_(0.5).bfloat16()
            else:
                b = torch.randn((K, N) if b_mn else (N, K), device=dev).mul_(0.05).bfloat16()
            kw = dict(a_mn=bool(a_mn), b_mn=bool(b_mn), bn_hint=o.get("bn", bn_hint))
            if o.get("conv"):
                kw.update(conv=(1, cb, ch, cw, cc, 3, 3, 1, 1), M=M, N=N, K=K)
            if o.get("wconv"):
                kw.update(conv=(2, cb, ch, cw, cc, 3, 3, 1, 1), M=M, N=N, K=K)
            if o.get("colsum"):
                st_ = torch.zeros(2, N, device=dev, dtype=torch.float64)
                kw.update(colsum=st_[0], colsumsq=st_[1])
            if o.get("stat"):
                st_ = torch.zeros(2, N, device=dev, dtype=torch.float64)
                kw.update(colsum=st_[0], colsumsq=st_[1], stat_x=torch.randn(M, N, device=dev).bfloat16(),
                          stat_mean=torch.zeros(N, device=dev), stat_scale=torch.ones(N, device=dev),
                          stat_shift=torch.zeros(N, device=dev), stat_relu=True)
            if o.get("acc"):
                kw.update(out=torch.zeros(M, N, device=dev), accumulate=True, split_k=-1)
            else:
                kw.update(out=torch.empty(M, N, device=dev, dtype=torch.bfloat16))
            if o.get("bias"):
                kw["bias"] = torch.randn(N, device=dev)
            if o.get("act"):
                kw["act"] = o["act"]
            if o.get("aux_out"):
                kw["aux_out"] = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            if o.get("dact"):
                kw["dact"] = o["dact"]
                kw["aux_in"] = torch.randn(M, N, device=dev).bfloat16()
            if o.get("residual"):
                kw["residual"] = torch.randn(M, N, device=dev).bfloat16()
            if o.get("drop"):
                kw.update(dropout_p=o["drop"], dropout_seed=1234)
            sets.append((a, b, kw))
        for a, b, kw in sets:
            ops.gemm(a, b, **kw)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(REP):
                a, b, kw = sets[i % ROT]
                ops.gemm(a, b, **kw)
        g.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / REP)
        t = sorted(ts)[len(ts) // 2]
        fl = 2.0 * M * N * K
        print(json.dumps({"name": name, "M": M, "N": N, "K": K, "us": round(t * 1e3, 2), "tflops": round(fl / t / 1e9, 1)}), flush=True)


if __name__ == "__main__":
    main()

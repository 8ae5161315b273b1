"""BatchNorm streaming kernels at the ResNet-50 layer shapes of the benchmark (B = 128): per-kernel time and achieved HBM
bandwidth, timed with CUDA events over rotating buffer sets larger than L2.  Usage: python tools/bench_bn.py [tag]
(MDHS_BN_REDUCE="RL,blocks_per_sm" selects the backward-reduce grid)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mdhs_b200  # noqa
from mdhs_b200 import ops

B = 128
SHAPES = [("stem 112x112x64", B * 112 * 112, 64, False), ("l1 56x56x64", B * 56 * 56, 64, False), ("l1 56x56x256", B * 56 * 56, 256, True),
          ("l2 28x28x128", B * 28 * 28, 128, False), ("l2 28x28x512", B * 28 * 28, 512, True),
          ("l3 14x14x256", B * 14 * 14, 256, False), ("l3 14x14x1024", B * 14 * 14, 1024, True),
          ("l4 7x7x512", B * 7 * 7, 512, False), ("l4 7x7x2048", B * 7 * 7, 2048, True)]
PEAK = 6530.0


def timeit(fn, sets, iters=20):
    for i in range(3):
        fn(sets[i % len(sets)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(sets[i % len(sets)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


rows_out = []
for name, rows, C, has_res in SHAPES:
    nbytes = rows * C * 2
    nsets = max(2, min(8, int(400e6 // (nbytes * 3)) + 1))
    sets = []
    for _ in range(nsets):
        x = (torch.randn(rows, C, device="cuda") * 1.5 + 0.3).bfloat16()
        dy = torch.randn(rows, C, device="cuda").bfloat16()
        res = torch.randn(rows, C, device="cuda").bfloat16() if has_res else None
        sets.append((x, dy, res))
    gamma, beta = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda")
    cs, cq = torch.zeros(C, device="cuda", dtype=torch.float64), torch.zeros(C, device="cuda", dtype=torch.float64)
    ops.col_stats(sets[0][0], cs, cq)
    y, mean, invstd, scale, shift = ops.bn_fwd(sets[0][0], cs, cq, gamma, beta, None, None, 0.1, 1e-5, residual=sets[0][2], relu=True)
    ys = [ops.bn_fwd(s[0], cs, cq, gamma, beta, None, None, 0.1, 1e-5, residual=s[2], relu=True)[0] if has_res else None for s in sets]
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    sums = torch.zeros(2, C, device="cuda", dtype=torch.float64)
    idx = {id(s): i for i, s in enumerate(sets)}

    t_fwd = timeit(lambda s: ops.bn_fwd(s[0], cs, cq, gamma, beta, None, None, 0.1, 1e-5, residual=s[2], relu=True), sets)
    t_bwd = timeit(lambda s: ops.bn_bwd(s[1], s[0], ys[idx[id(s)]], mean, invstd, gamma, dg, db, relu=True, want_dz=has_res,
                                        scale=scale, shift=shift), sets)
    t_apply = timeit(lambda s: ops.bn_bwd(s[1], s[0], ys[idx[id(s)]], mean, invstd, gamma, dg, db, relu=True, want_dz=has_res,
                                          scale=scale, shift=shift, sums=sums), sets)
    t_stats = timeit(lambda s: ops.col_stats(s[0], cs, cq), sets)
    n_in_f = 2 if has_res else 1
    b_fwd = nbytes * (n_in_f + 1)
    n_in_b = 3 if has_res else 2
    b_red = nbytes * n_in_b
    b_app = nbytes * (n_in_b + (2 if has_res else 1))
    t_red = t_bwd - t_apply
    rec = dict(shape=name, MB=round(nbytes / 1e6, 1), fwd_us=round(t_fwd, 1), fwd_frac=round(b_fwd / t_fwd / 1e3 / PEAK, 3),
               reduce_us=round(t_red, 1), reduce_frac=round(b_red / max(t_red, 1e-3) / 1e3 / PEAK, 3),
               apply_us=round(t_apply, 1), apply_frac=round(b_app / t_apply / 1e3 / PEAK, 3),
               col_stats_us=round(t_stats, 1), col_stats_frac=round(nbytes / t_stats / 1e3 / PEAK, 3))
    rows_out.append(rec)
    print(json.dumps(rec), flush=True)
    del sets, ys
    torch.cuda.empty_cache()
tag = sys.argv[1] if len(sys.argv) > 1 else "bn"
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"reduce_cfg": os.environ.get("MDHS_BN_REDUCE", "default(8,2)"), "rows": rows_out},
          open(os.path.join(ROOT, "gpurun_out", f"bench_bn_{tag}.json"), "w"), indent=1)

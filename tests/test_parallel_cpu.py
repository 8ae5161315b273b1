"""Host-side data-parallel logic on CPU: world_size-2 gloo run of GradSync (the same code path NCCL uses on the
GPUs), plus the flat-buffer bucket boundary."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mdhs_b200  # noqa: F401
    from mdhs_b200.parallel import GradSync
    n, split = 1000, 300
    torch.manual_seed(rank)
    g = torch.randn(n)
    mine = g.clone()
    sync = GradSync(g, split)
    sync.reduce_tail()                      # tail slice first (overlapped with the rest of backward on the GPUs)
    scale = sync.finish()
    others = []
    for r in range(world):
        torch.manual_seed(r)
        others.append(torch.randn(n))
    want = sum(others)
    ok = torch.allclose(g, want, atol=1e-6) and abs(scale - 1.0 / world) < 1e-12
    # without the early call everything goes out in finish()
    g2 = mine.clone()
    sync2 = GradSync(g2, split)
    sync2.finish()
    ok = ok and torch.allclose(g2, want, atol=1e-6)
    # per-stage buckets inside the head slice (ResNet layer4 / layer3 leave early), in any order, plus the tail
    g3 = mine.clone()
    sync3 = GradSync(g3, split)
    sync3.reduce_range(200, 290)
    sync3.reduce_tail()
    sync3.reduce_range(120, 200)
    sync3.finish()
    ok = ok and torch.allclose(g3, want, atol=1e-6)
    # bf16 buckets: every slice is cast into the bf16 communication buffer right before its all-reduce; the
    # reduced gradient is read from `sync.reduced`, the fp32 buffer keeps the local gradient; the first collective of a step
    # triggers the `on_first_reduce` callback exactly once (the trainer shrinks the persistent GEMM grids there)
    g4 = mine.clone()
    fired = []
    sync4 = GradSync(g4, split, comm_dtype=torch.bfloat16, on_first_reduce=lambda: fired.append(1))
    sync4.reduce_range(640, 1000)           # BERT layers 8..11 + fusion + head
    sync4.reduce_range(470, 640)            # BERT layers 4..7
    sync4.reduce_range(300, 470)            # embeddings + layers 0..3
    sync4.finish()
    want16 = sum(o.bfloat16().float() for o in others)
    ok = ok and sync4.reduced is not None and sync4.reduced.dtype == torch.bfloat16
    ok = ok and torch.allclose(sync4.reduced.float(), want16, atol=2e-2, rtol=1e-2) and torch.equal(g4, mine)
    ok = ok and len(fired) == 1
    sync4.reduce_tail()
    sync4.finish()
    ok = ok and len(fired) == 2
    # pipelined tail (Trainer overlap_comm="pipeline"): buckets go out in order, each is waited for individually
    g5 = mine.clone()
    sync5 = GradSync(g5, 0)
    bk = [(0, 100), (100, 400), (400, 1000)]
    for lo, hi in bk:
        sync5.reduce_range(lo, hi)
    for i, (lo, hi) in enumerate(bk):
        sync5.wait(i)
        ok = ok and torch.allclose(g5[lo:hi], want[lo:hi], atol=1e-6)
    sync5.finish()
    ok = ok and torch.allclose(g5, want, atol=1e-6)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_gradsync_gloo_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29600 + (os.getpid() % 200)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))


def test_split_offset_orders_text_encoder_after_image_encoder():
    import torch.nn as nn
    import mdhs_b200  # noqa: F401
    from mdhs_b200.parallel import split_offset

    class FakeStore:
        def __init__(self, mod):
            self.offsets, off = {}, 0
            for p in mod.parameters():
                self.offsets[id(p)] = off
                off += p.numel()

    class M(nn.Module):
        def __init__(self):
            super().__init__()
            self.image_encoder = nn.Linear(4, 4)
            self.text_encoder = nn.Linear(4, 2)
            self.head = nn.Linear(2, 2)

    m = M()
    st = FakeStore(m)
    assert split_offset(st, m.text_encoder) == 4 * 4 + 4
    assert split_offset(st, None) == 0
    from mdhs_b200.parallel import param_range
    assert param_range(st, m.text_encoder) == (20, 20 + 4 * 2 + 2)
    assert param_range(st, nn.ReLU()) == (0, 0)

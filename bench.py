#!/usr/bin/env python
"""bench.py -- training throughput of the multimodal hot path on B200 (BASELINE.json: train samples/sec).

  python bench.py --gpus N --steps K --warmup W              # our arm (one process per GPU under torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K ...    # the UNMODIFIED reference (oracle/_ref) on the host cores
  python bench.py --impl torch_gpu --steps K ...             # the UNMODIFIED reference, torch eager on the same B200
  python bench.py --config {1..5}                            # the other BASELINE.json configurations (default 2)

Workload at N=1 (default, --config 2) = BASELINE.json configs[1]: ResNet-50 + BERT-base, cross-attention fusion
(`fusion_type: basic`), MLP head, HAM 7 classes, per-GPU batch 128, 3x224x224 images, sequence length 64, bf16 tensor-core
math, label-smoothed CE, AdamW lr 2e-4 -- forward + backward + gradient all-reduce + optimizer step, synthetic data,
random-init weights.  One JSON line is printed by rank 0 (see the task contract for the keys).

The line also carries: `roofline` (dominant kernel = the tcgen05 GEMM: its own 2MNK FLOPs / its in-graph CUPTI time),
`cpu_baseline` (the reference on the host cores, bounded sample), `gpu_baseline` (the reference with torch eager on the
same B200: fp32 as shipped and bf16 autocast + channels_last -- the stock-GPU yardstick of SURVEY 2b / 8d), `e2e`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "train_samples_per_sec"
UNIT = "samples/s"

# BASELINE.json configs -> concrete shapes.  gflop = fwd+bwd GFLOP per sample (FlopCounterMode on the
# reference, 2 x MAC); config 4 = ConvNeXt-Tiny features (8.91 fwd) + BERT S=64 (10.87 fwd), x3.
CONFIGS = {
    1: dict(family="baseline", fusion="concat", classes=7, batch=32, seq=64, gflop=57.06,
            name="ResNet-50 + BERT-base, concat fusion, MLP head, HAM 7-class"),
    2: dict(family="baseline", fusion="basic", classes=7, batch=128, seq=64, gflop=57.50,
            name="ResNet-50 + BERT-base, basic (cross-attention) fusion, MLP head, HAM 7-class"),
    3: dict(family="baseline", fusion="multiscale", classes=6, batch=128, seq=64, gflop=59.37,
            name="ResNet-50 layer2/3/4 x BERT-base, multiscale cross-attention fusion, MLP head, Spine 6-class"),
    4: dict(family="connext_moe", fusion=None, classes=7, batch=128, seq=64, gflop=59.3,
            name="ConvNeXt-Tiny + BERT-base CLS -> MoE head (4 KAN experts, top-2), HAM 7-class"),
    5: dict(family="mibf", fusion=None, classes=6, batch=128, seq=256, gflop=154.80,
            name="MIBF-Net: ResNet-50 + BERT-base, IBFA attention, MP-Loss, Spine 6-class"),
}


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d.get("bf16_tflops_sustained", 1401.0), d.get("hbm_gbs", 6530.0), "measured"
    return 1590.0, 6650.0, "fallback"


def workload(args):
    cfg = dict(CONFIGS[args.config])
    if args.batch:
        cfg["batch"] = args.batch
    if args.seq:
        cfg["seq"] = args.seq
    if args.fusion and cfg["family"] == "baseline":
        cfg["fusion"] = args.fusion
        cfg["name"] = f"ResNet-50 + BERT-base, {args.fusion} fusion, MLP head, {cfg['classes']}-class"
        cfg["gflop"] = {("basic", 64): 57.50, ("basic", 128): 90.27, ("multiscale", 64): 59.37, ("multiscale", 128): 92.57,
                        ("concat", 64): 57.06}.get((args.fusion, cfg["seq"]), cfg["gflop"])
    return cfg


def config_dict(cfg, args, world, use_graph=True):
    """The `config` object of the JSON line -- identical for our arm and the reference arms (same workload)."""
    return {"workload": f"{cfg['name']}, train step (fwd+bwd+all-reduce+optimizer)", "baseline_config": args.config,
            "per_gpu_batch": cfg["batch"], "global_batch": cfg["batch"] * world, "seq_len": cfg["seq"], "image": "3x224x224",
            "parallelism": f"dp{world}", "cuda_graph": use_graph,
            "l2": "per-step working set (>10 GB of activations) far exceeds the 126 MB L2; no explicit flush"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def synthetic_batch(B, S, num_classes, seed=123, image_hw=224, unit_range=False):
    """Synthetic inputs of SURVEY.md section 8d (ImageNet-normalised-like randn images -- [0,1] `rand` for the MIBF / ConNexT
    families, which feed un-normalised pixels --, ids with CLS = 101, tail-padded mask, labels): bench.py's own copy, our arm
    does not touch anything under oracle/."""
    import torch
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    images = torch.rand(B, 3, image_hw, image_hw, generator=g) if unit_range else torch.randn(B, 3, image_hw, image_hw, generator=g)
    ids = torch.randint(0, 30522, (B, S), generator=g)
    ids[:, 0] = 101
    lens = torch.randint(min(8, S), S + 1, (B,), generator=g)
    mask = (torch.arange(S)[None, :] < lens[:, None]).long()
    labels = torch.randint(0, num_classes, (B,), generator=g)
    return images, ids, mask, labels


# ----------------------------------------------------------------------------------------------
# reference arms: the unmodified reference (oracle/_ref, vendored verbatim by oracle/make_ref.py) through its own public
# API and its own training loop body (scripts/train.py:364-387: CrossEntropyLoss(label_smoothing) + AdamW; MIBF:
# mibf_net/train_resnet.py:21-41).  Falls back to the oracle port (plain-PyTorch restatement) when oracle/_ref is absent.
# ----------------------------------------------------------------------------------------------
def reference_step_fn(cfg, device, batch, mode="fp32"):
    """Returns (step callable -> loss tensor, kind).  mode: "fp32" (as shipped) | "bf16_cl" (autocast + channels_last)."""
    import torch
    import torch.nn as nn
    from oracle import make_ref
    from refutil import bert_dir, quiet
    images, ids, mask, labels = [t.to(device) for t in batch]
    autocast = mode == "bf16_cl"
    if autocast:
        images = images.contiguous(memory_format=torch.channels_last)
    torch.manual_seed(0)
    fam = cfg["family"]
    if make_ref.available() or os.path.isdir("/root/reference"):
        kind = "reference"
        if fam == "baseline":
            from refutil import build_reference_model
            model = build_reference_model(fusion=cfg["fusion"], head="mlp", num_classes=cfg["classes"])
            crit = nn.CrossEntropyLoss(label_smoothing=0.02)

            def fwd():
                return crit(model(images, ids, mask), labels)
        elif fam == "mibf":
            import torchvision
            from refutil import REF_ROOT
            if REF_ROOT not in sys.path:
                sys.path.insert(0, REF_ROOT)
            orig = torchvision.models.resnet50
            torchvision.models.resnet50 = lambda *a, **k: orig(weights=None)   # offline: random init instead of the download
            try:
                with quiet():
                    from mibf_net.model_resnet import Resnet50WithOurs
                    model = Resnet50WithOurs(num_labels=cfg["classes"], bert_path=bert_dir())
            finally:
                torchvision.models.resnet50 = orig

            def fwd():
                out = model({"transformed_image": images, "input_ids": ids, "attention_mask": mask})
                return model.cal_loss(out, labels)
        else:
            kind = None
    else:
        kind = None
    if kind is None:
        # oracle port (configs without an importable reference composite, or no vendored reference on this box)
        kind = "port"
        from oracle import port, weights
        import mdhs_b200  # noqa: F401  (state_dict template only: keys / shapes; never used for compute here)
        if fam == "baseline":
            from refutil import build_ours
            tmpl = build_ours(fusion=cfg["fusion"], head="mlp", num_classes=cfg["classes"]).state_dict()
        elif fam == "mibf":
            from mdhs_b200.mibf_net.model_resnet import Resnet50WithOurs as Ours
            with quiet():
                tmpl = Ours(num_labels=cfg["classes"], bert_path=bert_dir(), pretrained=False).state_dict()
        else:
            raise RuntimeError("no reference composite for this configuration (pl_model_MOE2 needs pytorch_lightning)")
        sd = weights.synth_state_dict(tmpl, seed=0)
        params = {k: v.clone().to(device).requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
        state = {k: v.to(device) for k, v in sd.items()}
        state.update(params)

        class _Holder(nn.Module):
            def __init__(self):
                super().__init__()
                self.plist = nn.ParameterList([nn.Parameter(p) for p in params.values()])
        model = _Holder()
        for (k, _), p in zip(params.items(), model.plist):
            state[k] = p

        if fam == "baseline":
            def fwd():
                logits = port.model_forward(state, images, ids, mask, fusion=cfg["fusion"], head="mlp", training_bn=True)
                return port.ce_label_smoothing(logits, labels, label_smoothing=0.02)
        else:
            def fwd():
                out = port.mibf_forward(state, images, ids, mask, training_bn=True)
                return port.mp_loss(out["image"], out["text"], out["image_text"], labels)
    model = model.to(device).train()
    if autocast:
        model = model.to(memory_format=torch.channels_last)
    if fam == "mibf":
        opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9)   # mibf_net/train_resnet.py:136-139
    else:
        opt = torch.optim.AdamW(model.parameters(), lr=2e-4)               # scripts/train.py:283-299, config.yml:80-84

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast(device_type=device.type, dtype=torch.bfloat16, enabled=autocast):
            loss = fwd()
        loss.backward()
        opt.step()
        return loss.detach()
    return step, kind


def cpu_reference_samples_per_sec(cfg, batch_size, steps, warmup):
    import torch
    torch.set_num_threads(os.cpu_count())
    unit = cfg["family"] != "baseline"
    batch = synthetic_batch(batch_size, cfg["seq"], cfg["classes"], unit_range=unit)
    step, kind = reference_step_fn(cfg, torch.device("cpu"), batch)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return batch_size * len(times) / total, total / len(times), kind


def torch_gpu_samples_per_sec(cfg, dev, steps, warmup, mode):
    """The reference modules with torch eager (cuDNN / cuBLAS / ATen) on the same B200; CUDA-event timed."""
    import torch
    unit = cfg["family"] != "baseline"
    batch = synthetic_batch(cfg["batch"], cfg["seq"], cfg["classes"], unit_range=unit)
    step, kind = reference_step_fn(cfg, dev, batch, mode)
    for _ in range(max(warmup, 3)):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": round(cfg["batch"] / (ms / 1e3), 1), "unit": UNIT, "ms_per_step": round(ms, 3), "steps": steps,
            "kind": kind, "final_loss": round(float(loss.item()), 4)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = workload(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    batch = min(args.ref_batch, cfg["batch"])
    sps, sec, kind = cpu_reference_samples_per_sec(cfg, batch, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(sps, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(sec * 1e3, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(cfg, args, world),
        "note": f"the reference's own CPU PyTorch path ({'unmodified reference modules from oracle/_ref' if kind == 'reference' else 'oracle port'}"
                f", fp32, {os.cpu_count()} host threads); each step is a bounded sample of {batch} of the {cfg['batch']} samples",
        "cpu_baseline": {"value": round(sps, 3), "unit": UNIT, "cores": os.cpu_count(), "kind": kind,
                         "sample": f"{args.steps} steps of batch {batch} after {args.warmup} warm-up"},
        "e2e": {"value": round(sps, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_torch_gpu(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cfg = workload(args)
    dev = torch.device("cuda", 0)
    res = {}
    for mode in ("fp32", "bf16_cl"):
        try:
            res[mode] = torch_gpu_samples_per_sec(cfg, dev, args.steps, args.warmup, mode)
        except Exception as e:
            res[mode] = {"error": f"{type(e).__name__}: {e}"}
        torch.cuda.empty_cache()
    best = max((r["value"] for r in res.values() if "value" in r), default=None)
    line = {"impl": "torch_gpu", "metric": METRIC, "value": best, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 / bf16-autocast", "data": "synthetic", "config": config_dict(cfg, args, 1, use_graph=False),
            "gpu_baseline": {"fp32_as_shipped": res["fp32"], "bf16_autocast_channels_last": res["bf16_cl"]},
            "note": "the reference's modules with torch eager (cuDNN/cuBLAS/ATen) on one B200; `value` = the faster of the two"}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def build_ours_model(cfg, dev, dp_kw=None):
    """Our B200 model + Trainer for one BASELINE configuration.  dp_kw: data-parallel knobs of the Trainer."""
    dp_kw = dp_kw or {}
    import torch
    import mdhs_b200
    from mdhs_b200.train import Trainer, mibf_forward_loss
    from mdhs_b200 import functional as Fm
    from refutil import bert_dir, quiet
    torch.manual_seed(0)
    fam = cfg["family"]
    with quiet():
        if fam == "baseline":
            model = mdhs_b200.MultimodalBaselineModel(num_classes=cfg["classes"], hidden_dim=256, dropout=0.2, pretrained_image=False,
                                                      image_weights_path=None, text_model_name=bert_dir(), num_heads=8,
                                                      image_backbone="resnet50", classifier_type="mlp", fusion_type=cfg["fusion"])
            model = model.to(dev)
            trainer = Trainer(model, optimizer="adamw", lr=2e-4, label_smoothing=0.02, **dp_kw)
        elif fam == "mibf":
            from mdhs_b200.mibf_net.model_resnet import Resnet50WithOurs
            model = Resnet50WithOurs(num_labels=cfg["classes"], bert_path=bert_dir(), pretrained=False).to(dev)
            trainer = Trainer(model, optimizer="sgd", lr=1e-3, momentum=0.9, forward_loss=mibf_forward_loss)
        else:
            from mdhs_b200.connext.ourmodel import ConvNeXtMoEClassifier
            model = ConvNeXtMoEClassifier(num_labels=cfg["classes"], variant="tiny", use_text=True, bert_path=bert_dir()).to(dev)

            def moe_forward_loss(m, images, ids, mask, labels):
                logits, aux = m({"transformed_image": images, "input_ids": ids, "attention_mask": mask})
                return Fm.cross_entropy(logits, labels) + aux, logits
            trainer = Trainer(model, optimizer="adamw", lr=2e-4, forward_loss=moe_forward_loss)
    return model, trainer


def run_ours(args):
    # rank 0's stdout must carry exactly ONE JSON line, but native libraries write there too (NCCL prints its version banner
    # with printf): park the real stdout, point fd 1 at stderr for the whole run, and emit the line on the parked descriptor
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # rank 0's stdout carries exactly ONE JSON line: NCCL's own banner / debug output (NCCL_DEBUG=VERSION prints
        # "NCCL version ..." to stdout) goes to stderr instead
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        # collectives are captured into the step's CUDA graph: the NCCL watchdog's async error handling must not
        # poll events of a capturing stream (torch CUDA-graphs notes)
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
        dist.init_process_group("nccl", device_id=dev)
    import mdhs_b200  # noqa: F401
    from mdhs_b200 import _lib, ops
    from refutil import bert_dir

    cfg = workload(args)
    B, S, HW, C = cfg["batch"], cfg["seq"], 224, cfg["classes"]
    if rank == 0:
        bert_dir()
    if world > 1:
        dist.barrier()
    dp_kw = dict(comm_dtype=args.comm_dtype, bert_bucket_layers=args.bert_bucket_layers, overlap_comm=args.overlap_comm)
    if args.sm_reserve >= 0:
        dp_kw["sm_reserve"] = args.sm_reserve
    model, trainer = build_ours_model(cfg, dev, dp_kw)
    images, ids, mask, labels = synthetic_batch(B, S, C, seed=123 + rank, image_hw=HW, unit_range=cfg["family"] != "baseline")
    h_in = [t.pin_memory() for t in (images, ids, mask, labels)]
    d_in = [t.to(dev, non_blocking=True) for t in h_in]
    torch.cuda.synchronize()

    use_graph = not args.no_graph

    def _mark(msg):
        if args.verbose:
            print(f"[bench rank {rank}] {msg}", file=sys.stderr, flush=True)
    _mark("model built")
    launches_per_step = None
    if use_graph:
        try:
            trainer.capture(*d_in, warmup=3)
            c0 = _lib.launch_count()
        except Exception as e:  # e.g. a collective that cannot be captured: fall back to eager launches
            if rank == 0:
                print(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); running eager", file=sys.stderr)
            use_graph = False
            torch.cuda.synchronize()

    def one_step():
        if use_graph:
            return trainer.replay()
        return trainer.step(*d_in)

    _mark(f"first step done (graph={use_graph})")
    for _ in range(max(args.warmup, 3)):
        one_step()
    torch.cuda.synchronize()
    _mark("warm-up done")
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        loss, _ = one_step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    _mark(f"timed region done: {ms / args.steps:.2f} ms/step")
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    value = world * B * args.steps / (ms / 1e3)
    final_loss = float(loss.item())

    # ---- end-to-end: pinned host inputs copied every step (double-buffered on a copy stream), loss read back
    copy_stream = torch.cuda.Stream()
    bufs = [[torch.empty_like(t, device=dev) for t in h_in] for _ in range(2)]
    h_loss = torch.empty(1, dtype=torch.float32).pin_memory()

    def upload(slot):
        with torch.cuda.stream(copy_stream):
            for dst, src in zip(bufs[slot], h_in):
                dst.copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    def e2e_loop(n):
        ev = upload(0)
        for i in range(n):
            slot = i & 1
            torch.cuda.current_stream().wait_event(ev)
            if i + 1 < n:
                ev = upload(slot ^ 1)
            if use_graph:
                l, _ = trainer.replay(*bufs[slot])
            else:
                l, _ = trainer.step(*bufs[slot])
            h_loss.copy_(l.view(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()   # the host consumes the loss every step
    e2e_loop(2)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_loop(args.steps)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / e2e_s.item()
    h2d = sum(t.numel() * t.element_size() for t in h_in)

    # ---- roofline of the dominant kernel (the tcgen05 GEMM).  FLOPs: the kernel's OWN work, sum of 2*M*N*K over the GEMM
    # launches of one step (recorded by wrapping the C-ABI call in one eager step).  Time: CUPTI durations of exactly those
    # kernels inside the graph-replayed step (warm, no launch gaps); falls back to CUDA-event pairs around eager launches.
    roof = None
    if rank == 0 and world == 1:
        recs = []
        real_gemm = ops.gemm

        def timed_gemm(a, b, **kw):
            s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_.record()
            out = real_gemm(a, b, **kw)
            e_.record()
            a_mn, b_mn = kw.get("a_mn", False), kw.get("b_mn", False)
            M = kw.get("M") or (a.shape[1] if a_mn else a.shape[0])
            N = kw.get("N") or (b.shape[1] if b_mn else b.shape[0])
            K = kw.get("K") or (a.shape[0] if a_mn else a.shape[1])
            recs.append((2.0 * M * N * K, s_, e_, (M, N, K, int(a_mn), int(b_mn), int(bool(kw.get("accumulate"))), kw.get("split_k", 1),
                                                   int(kw.get("conv") is not None))))
            return out
        ops.gemm = timed_gemm
        c0 = _lib.launch_count()
        try:
            s_all, e_all = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            trainer.step(*d_in)          # allocator warm-up on this stream (the captured graph owns its own pool)
            recs.clear()
            torch.cuda.synchronize()
            c0 = _lib.launch_count()
            torch.cuda._sleep(int(60e6))  # ~30 ms of GPU spin: the host runs ahead, so event pairs bracket kernels only
            s_all.record()
            trainer.step(*d_in)
            e_all.record()
            torch.cuda.synchronize()
            launches_per_step = _lib.launch_count() - c0
        finally:
            ops.gemm = real_gemm
        gemm_us = other = None
        top = None
        excl = None      # kernels timed one at a time (single-stream replay) when the step itself overlaps two streams

        def cupti_totals(reps=2):
            import collections
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                for _ in range(reps):
                    trainer.replay()
                torch.cuda.synchronize()
            agg = collections.defaultdict(lambda: [0, 0.0])
            for e in prof.events():
                if e.device_type == torch.autograd.DeviceType.CUDA and e.name and "Memcpy" not in e.name and "Memset" not in e.name:
                    k = e.name.replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "").split("(")[0]
                    k = k.split("<")[0]
                    agg[k][0] += 1
                    agg[k][1] += e.time_range.end - e.time_range.start
            g_us = agg["gemm_tc_kernel"][1] / reps if "gemm_tc_kernel" in agg else None
            busy_ = sum(v[1] for v in agg.values()) / reps
            top_ = [[k, round(v[0] / reps, 1), round(v[1] / reps, 1)] for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]]
            return g_us, busy_, top_
        if use_graph:
            try:
                gemm_us, other, top = cupti_totals()
                from mdhs_b200 import runtime as _rt
                if _rt.DUAL_STREAM and world == 1 and gemm_us:
                    # the timed step runs the two encoders on two streams: concurrent kernels share SMs and their CUPTI
                    # durations are not exclusive.  For the kernel's roofline, replay the SAME step captured on one stream.
                    _rt.DUAL_STREAM = False
                    try:
                        trainer.release_graph()
                        trainer.capture(*d_in, warmup=1)
                        g1, b1, t1 = cupti_totals()
                        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        ev0.record()
                        for _ in range(5):
                            trainer.replay()
                        ev1.record()
                        torch.cuda.synchronize()
                        excl = {"gemm_us": g1, "busy_us": b1, "top": t1, "step_ms": ev0.elapsed_time(ev1) / 5}
                    finally:
                        _rt.DUAL_STREAM = True
                        trainer.release_graph()
                        trainer.capture(*d_in, warmup=1)     # back to the step that was timed (timeline / e2e below)
            except Exception as e:
                print(f"[bench] CUPTI timeline unavailable ({type(e).__name__}: {e}); using eager CUDA-event pairs", file=sys.stderr)
        if recs:
            gflop = sum(r[0] for r in recs) / 1e9
            gms_eager = sum(r[1].elapsed_time(r[2]) for r in recs)
            gms_step = gemm_us / 1e3 if gemm_us else gms_eager          # durations inside the timed step
            gms = excl["gemm_us"] / 1e3 if excl else gms_step           # exclusive durations (= gms_step on one stream)
            peak, hbm, src = _peaks()
            ach = gflop / gms  # GFLOP/ms == TFLOP/s
            if args.dump_gemms:
                import collections
                agg2 = collections.defaultdict(lambda: [0, 0.0, 0.0])
                for r in recs:
                    a_ = agg2[r[3]]
                    a_[0] += 1
                    a_[1] += r[1].elapsed_time(r[2])
                    a_[2] += r[0]
                os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
                with open(os.path.join(ROOT, "gpurun_out", f"gemm_shapes_c{args.config}.txt"), "w") as fh:
                    fh.write("M N K a_mn b_mn acc split conv | count total_ms TFLOP/s   (eager CUDA-event pairs)\n")
                    for k_, (c_, ms_, fl_) in sorted(agg2.items(), key=lambda kv: -kv[1][1]):
                        fh.write(f"{k_} | {c_} {ms_:.3f} {fl_ / ms_ / 1e9:.1f}\n")
            # DRAM traffic of the same kernels: from the committed ncu capture of this command (profiles/), per launch
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "r02_gemm_dram_traffic.json")
            if os.path.exists(tpath) and args.config == 2:
                try:
                    tj = json.load(open(tpath))
                    traffic = round(tj["dram_bytes_per_step"] / tj["launches_per_step"])
                except Exception:
                    traffic = None
            roof = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 bf16 GEMM: all Linear / conv contractions)",
                    "achieved": round(ach, 1), "peak": peak, "unit": "TFLOP/s", "frac": round(ach / peak, 4), "traffic": traffic,
                    "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read+write, mean over the step's GEMM launches; "
                                    "profiles/r02_gemm_dram_traffic.json)",
                    "peak_source": f"{src} bf16_tflops_sustained", "launches_per_step": len(recs),
                    "timing": ("CUPTI kernel durations inside a single-stream graph replay of the same step (the timed step runs the "
                               "two encoders on two streams; durations of concurrent kernels are not exclusive: "
                               "gemm_ms_per_step_overlapped)" if excl else
                               "CUPTI kernel durations inside the graph-replayed step" if gemm_us else "CUDA-event pairs, eager step"),
                    "gemm_ms_per_step": round(gms, 3), "gemm_ms_per_step_eager_events": round(gms_eager, 3),
                    "gemm_ms_per_step_overlapped": round(gms_step, 3) if excl else None,
                    "single_stream_ms_per_step": round(excl["step_ms"], 3) if excl else None,
                    "gemm_gflop_per_step": round(gflop, 1),
                    "flops_per_launch": round(gflop * 1e9 / len(recs)), "us_per_launch": round(gms * 1e3 / len(recs), 2),
                    "gemm_share_of_step": round(gms / (excl["step_ms"] if excl else ms / args.steps), 3),
                    "kernel_busy_ms_per_step": round((excl["busy_us"] if excl else other) / 1e3, 3) if other else None,
                    "top_kernels_us": excl["top"] if excl else top,
                    "top_kernels_us_overlapped": top if excl else None,
                    "model_flops_frac": round(value / world * cfg["gflop"] / 1e3 / peak, 4)}
    # ---- optional kernel timeline of ONE graph-replayed step (CUPTI through torch.profiler), every rank's view of rank 0:
    # name / stream / start / duration of every kernel, so that exposed collectives and stretched kernels can be read off
    if args.timeline and rank == 0 and use_graph:
        try:
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                trainer.replay()
                torch.cuda.synchronize()
            import re as _re

            def _kname(n):
                n = n.replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
                m_ = _re.match(r"[A-Za-z_0-9:]+(<[^(]*>)?", n)
                return (m_.group(0) if m_ else n)[:96]
            rows_ = []
            try:        # kineto events carry the stream id
                for e in prof.profiler.kineto_results.events():
                    if str(e.device_type()).endswith("CUDA") and e.name():
                        rows_.append({"name": _kname(e.name()), "start_us": e.start_ns() / 1e3, "dur_us": round(e.duration_ns() / 1e3, 2),
                                      "stream": int(e.device_resource_id())})
            except Exception:
                rows_ = [{"name": _kname(e.name), "start_us": e.time_range.start, "dur_us": round(e.time_range.end - e.time_range.start, 2),
                          "stream": -1} for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.name]
            rows_.sort(key=lambda r: r["start_us"])
            t0 = rows_[0]["start_us"]
            for r in rows_:
                r["start_us"] = round(r["start_us"] - t0, 2)
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            json.dump({"world": world, "ms_per_step": ms / args.steps, "kernels": rows_},
                      open(os.path.join(ROOT, "gpurun_out", args.timeline), "w"))
        except Exception as e:
            print(f"[bench] timeline failed: {type(e).__name__}: {e}", file=sys.stderr)
    elif args.timeline and use_graph:
        trainer.replay()        # keep the ranks in lock-step with rank 0's profiled replay (the step contains collectives)
        torch.cuda.synchronize()
    if launches_per_step is None:
        c0 = _lib.launch_count()
        trainer.step(*d_in)
        torch.cuda.synchronize()
        launches_per_step = _lib.launch_count() - c0
    if world > 1:
        dist.barrier()

    # ---- inference throughput (BASELINE.json's second metric): eval-mode forward of the same batch, one CUDA graph
    infer = None
    if not args.no_inference and cfg["family"] == "baseline":
        try:
            model.eval()
            s_in = [t.clone() for t in d_in[:3]]
            with torch.no_grad():
                for _ in range(3):
                    model(*s_in)
                torch.cuda.synchronize()
                use_g = not args.no_graph
                g_inf = None
                if use_g:
                    try:
                        g_inf = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g_inf):
                            logits_inf = model(*s_in)  # noqa: F841
                    except Exception:
                        g_inf = None
                        torch.cuda.synchronize()
                n_inf = max(args.steps, 10)
                for _ in range(3):
                    g_inf.replay() if g_inf is not None else model(*s_in)
                torch.cuda.synchronize()
                i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                i0.record()
                for _ in range(n_inf):
                    g_inf.replay() if g_inf is not None else model(*s_in)
                i1.record()
                torch.cuda.synchronize()
            ti = torch.tensor([i0.elapsed_time(i1) / n_inf], device=dev)
            if world > 1:
                dist.all_reduce(ti, op=dist.ReduceOp.MAX)
            infer = {"value": round(world * B / (ti.item() / 1e3), 1), "unit": UNIT, "ms_per_batch": round(ti.item(), 3),
                     "per_gpu_batch": B, "cuda_graph": g_inf is not None, "mode": "eval forward, logits on device"}
            g_inf = None
            model.train()
        except Exception as e:  # the training numbers above stay valid
            infer = {"error": f"{type(e).__name__}: {e}"}
    if world > 1:
        dist.barrier()

    # ---- baselines measured in the same run (rank 0, N = 1 only): the reference with torch eager on this B200, and on the
    # host cores.  Both are REPORTED baselines; neither touches our package.
    gpu_base = cpu = None
    if rank == 0 and world == 1 and not args.no_gpu_baseline and cfg["family"] in ("baseline", "mibf"):
        trainer.release_graph()
        torch.cuda.empty_cache()
        gpu_base = {}
        for mode, key in (("fp32", "fp32_as_shipped"), ("bf16_cl", "bf16_autocast_channels_last")):
            try:
                gpu_base[key] = torch_gpu_samples_per_sec(cfg, dev, 5, 3, mode)
            except Exception as e:
                gpu_base[key] = {"error": f"{type(e).__name__}: {e}"}
            torch.cuda.empty_cache()
        gpu_base["note"] = ("the reference's own modules, torch eager (cuDNN/cuBLAS/ATen) on the same B200, same batch / "
                            "optimizer / loss; 5 timed steps after 3 warm-up, CUDA events")
    if rank == 0 and world == 1 and not args.no_cpu_baseline and cfg["family"] in ("baseline", "mibf"):
        cb = min(32, B)
        sps, sec, kind = cpu_reference_samples_per_sec(cfg, cb, 2, 1)
        cpu = {"value": round(sps, 3), "unit": UNIT, "cores": os.cpu_count(), "kind": kind,
               "sample": f"2 steps of a {cb}-sample batch of this workload (S={S}, 224x224, fwd+bwd+optimizer) after 1 warm-up, fp32, "
                         f"{'unmodified reference modules (oracle/_ref)' if kind == 'reference' else 'oracle port'}"}
    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": config_dict(cfg, args, world, use_graph),
            "clocks": clocks, "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches_per_step) * args.steps, "gpu_launches_per_step": int(launches_per_step),
            "roofline": roof, "cpu_baseline": cpu, "gpu_baseline": gpu_base, "inference": infer, "final_loss": round(final_loss, 4),
        }
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        # the captured step holds NCCL work inside a CUDA graph; tearing the communicator down under it can block, so
        # drop the graph first, drain the device, and leave without the (optional) communicator destruction
        trainer.release_graph()
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch_gpu"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json configuration (1-based)")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the configuration's)")
    ap.add_argument("--seq", type=int, default=0, help="sequence length (default: the configuration's)")
    ap.add_argument("--fusion", default=None, choices=["basic", "multiscale", "concat"],
                    help="override the fusion_type of a MultimodalBaselineModel configuration")
    ap.add_argument("--ref-batch", type=int, default=32, help="samples per CPU step of the reference arm (bounded sample)")
    ap.add_argument("--comm-dtype", default="bf16", choices=["bf16", "fp32"], help="dtype of the gradient buckets on the wire (N > 1)")
    ap.add_argument("--sm-reserve", type=int, default=-1, help="SMs the persistent GEMM grids leave to NCCL while a bucket is in "
                                                               "flight (default: MDHS_SM_RESERVE or 16)")
    ap.add_argument("--bert-bucket-layers", type=int, default=0, help="BERT layers per early gradient bucket (0 = one bucket)")
    ap.add_argument("--overlap-comm", default="backward", choices=["pipeline", "backward", "none"],
                    help="N > 1: gradient buckets sent during backward (default), pipelined with the fused optimizer after "
                         "backward, or one all-reduce before the optimizer")
    ap.add_argument("--timeline", default=None, help="write the kernel timeline of one replayed step to gpurun_out/<name>")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-inference", action="store_true", help="skip the eval-forward throughput measurement")
    ap.add_argument("--graph-multi-gpu", action="store_true", help="(default now; kept for old command lines)")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--dump-gemms", action="store_true", help="write per-shape GEMM timings to gpurun_out/gemm_shapes_c<N>.txt")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "torch_gpu":
        run_torch_gpu(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

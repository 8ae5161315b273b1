#!/usr/bin/env python
"""bench.py -- training throughput of the multimodal hot path on B200 (BASELINE.json: train samples/sec).

  python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on the host cores

Workload at N=1 = BASELINE.json configs[1]: ResNet-50 + BERT-base, cross-attention fusion (`fusion_type: basic`),
MLP head, HAM 7 classes, per-GPU batch 128, 3x224x224 images, sequence length 64, bf16 tensor-core math,
label-smoothed CE, AdamW lr 2e-4 -- forward + backward + gradient all-reduce + optimizer step, synthetic data,
random-init weights.  One JSON line is printed by rank 0 (see the task contract for the keys).
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
FWD_BWD_GFLOP_PER_SAMPLE = 57.50   # ResNet-50 + BERT `basic` fusion, S=64 (FlopCounterMode, 2xMAC)
# SURVEY.md section 8d table (fwd+bwd GFLOP per sample) for the other --fusion / --seq combinations
_GFLOP = {("basic", 64): 57.50, ("basic", 128): 90.27, ("multiscale", 64): 59.37, ("multiscale", 128): 92.57, ("concat", 64): 57.06}


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d.get("bf16_tflops_sustained", 1401.0), d.get("hbm_gbs", 6530.0), "measured"
    return 1590.0, 6650.0, "fallback"


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


# ----------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port (plain PyTorch, fp32) on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_train_samples_per_sec(batch, steps, warmup, seq=64, hw=224):
    import torch
    from oracle import port, weights
    from refutil import build_ours  # only used for the state_dict template (keys/shapes), never for compute
    torch.set_num_threads(os.cpu_count())
    tmpl = build_ours(fusion="basic", head="mlp").state_dict()
    sd = weights.synth_state_dict(tmpl, seed=0)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
    state = dict(sd)
    state.update(params)
    opt = torch.optim.AdamW(list(params.values()), lr=2e-4)
    images, ids, mask, labels = weights.synthetic_batch(batch, seq, 7, image_hw=hw)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        logits = port.model_forward(state, images, ids, mask, fusion="basic", head="mlp", training_bn=True)
        loss = port.ce_label_smoothing(logits, labels, label_smoothing=0.02)
        loss.backward()
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return batch * len(times) / total, total / len(times)


def synthetic_batch(B, S, num_classes, seed=123, image_hw=224):
    """Synthetic inputs of SURVEY.md section 8d (ImageNet-normalised-like randn images, ids with CLS = 101, tail-padded mask,
    labels) -- bench.py's own copy: our arm does not touch anything under oracle/."""
    import torch
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    images = torch.randn(B, 3, image_hw, image_hw, generator=g)
    ids = torch.randint(0, 30522, (B, S), generator=g)
    ids[:, 0] = 101
    lens = torch.randint(min(8, S), S + 1, (B,), generator=g)
    mask = (torch.arange(S)[None, :] < lens[:, None]).long()
    labels = torch.randint(0, num_classes, (B,), generator=g)
    return images, ids, mask, labels


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = args.ref_batch
    sps, sec = cpu_train_samples_per_sec(batch, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(sps, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(sec * 1e3, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ResNet-50 + BERT-base, basic (cross-attention) fusion, MLP head, HAM 7-class, 3x224x224, seq 64, "
                               "fwd+bwd+AdamW", "per_step_batch": batch,
                   "note": "reference's own CPU PyTorch path (oracle port of the pure-Python reference), bounded sample per step"},
        "cpu_baseline": {"value": round(sps, 3), "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                         "sample": f"{args.steps} steps of batch {batch} after {args.warmup} warm-up"},
        "e2e": {"value": round(sps, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
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
    import mdhs_b200
    from mdhs_b200 import _lib, ops
    from mdhs_b200.train import Trainer
    from refutil import bert_dir, quiet   # tests/refutil.py: builds the local random-init bert-base directory

    B, S, HW, C = args.batch, args.seq, 224, 7
    if rank == 0:
        bert_dir()
    if world > 1:
        dist.barrier()
    torch.manual_seed(0)
    with quiet():
        model = mdhs_b200.MultimodalBaselineModel(num_classes=C, hidden_dim=256, dropout=0.2, pretrained_image=False,
                                                  image_weights_path=None, text_model_name=bert_dir(), num_heads=8,
                                                  image_backbone="resnet50", classifier_type="mlp", fusion_type=args.fusion)
    model = model.to(dev)
    trainer = Trainer(model, optimizer="adamw", lr=2e-4, label_smoothing=0.02)
    images, ids, mask, labels = synthetic_batch(B, S, C, seed=123 + rank, image_hw=HW)
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
            c0 = _lib.launch_count()
            trainer.capture(*d_in, warmup=3)
            # warm-up (3 eager) + 1 captured pass were recorded by the host-side counter
            launches_per_step = (_lib.launch_count() - c0) // 4
        except Exception as e:  # e.g. a collective that cannot be captured: fall back to eager launches
            if rank == 0:
                print(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); running eager", file=sys.stderr)
            use_graph = False
            torch.cuda.synchronize()
    if not use_graph:
        c0 = _lib.launch_count()
        trainer.step(*d_in)
        launches_per_step = _lib.launch_count() - c0

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

    # ---- roofline of the dominant kernel (the tcgen05 GEMM): one instrumented eager step, CUDA events per launch
    roof = None
    if rank == 0:
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
            recs.append((2.0 * M * N * K, s_, e_, (M, N, K, int(a_mn), int(b_mn), int(bool(kw.get("accumulate"))), kw.get("split_k", 1))))
            return out
        ops.gemm = timed_gemm
        try:
            s_all, e_all = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            if world == 1:
                trainer.step(*d_in)          # allocator warm-up on this stream (the captured graph owns its own pool)
                recs.clear()
                torch.cuda.synchronize()
                torch.cuda._sleep(int(60e6))  # ~30 ms of GPU spin: the host runs ahead, so event pairs bracket kernels only
                s_all.record()
                trainer.step(*d_in)
                e_all.record()
                torch.cuda.synchronize()
        finally:
            ops.gemm = real_gemm
        if recs:
            gflop = sum(r[0] for r in recs) / 1e9
            gms = sum(r[1].elapsed_time(r[2]) for r in recs)
            peak, hbm, src = _peaks()
            ach = gflop / gms  # GFLOP/ms == TFLOP/s
            if args.dump_gemms:
                import collections
                agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
                for r in recs:
                    a_ = agg[r[3]]
                    a_[0] += 1
                    a_[1] += r[1].elapsed_time(r[2])
                    a_[2] += r[0]
                os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
                with open(os.path.join(ROOT, "gpurun_out", "gemm_shapes.txt"), "w") as fh:
                    fh.write("M N K a_mn b_mn acc split | count total_ms TFLOP/s\n")
                    for k_, (c_, ms_, fl_) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                        fh.write(f"{k_} | {c_} {ms_:.3f} {fl_ / ms_ / 1e9:.1f}\n")
            roof = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 bf16 GEMM: all Linear / conv contractions)",
                    "achieved": round(ach, 1), "peak": peak, "unit": "TFLOP/s", "frac": round(ach / peak, 4), "traffic": None,
                    "peak_source": f"{src} bf16_tflops_sustained", "launches_per_step": len(recs),
                    "gemm_ms_per_step": round(gms, 3), "gemm_gflop_per_step": round(gflop, 1),
                    "gemm_share_of_eager_step": round(gms / s_all.elapsed_time(e_all), 3),
                    "model_flops_frac": round(value / world * _GFLOP.get((args.fusion, S), FWD_BWD_GFLOP_PER_SAMPLE) / 1e3 / peak, 4)}
    if world > 1:
        dist.barrier()

    # ---- inference throughput (BASELINE.json's second metric): eval-mode forward of the same batch, one CUDA graph
    infer = None
    if not args.no_inference:
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
                            logits_inf = model(*s_in)
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

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sps, sec = cpu_train_samples_per_sec(32, 2, 1)
        cpu = {"value": round(sps, 3), "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": "2 steps of batch 32 (BASELINE config 1 shape: S=64, 224x224, fwd+bwd+AdamW) after 1 warm-up, fp32 oracle port"}
    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"ResNet-50 + BERT-base, {args.fusion} (cross-attention) fusion, MLP head, HAM 7-class, train step "
                                   "(fwd+bwd+all-reduce+AdamW)", "per_gpu_batch": B, "global_batch": B * world, "seq_len": S,
                       "image": "3x224x224", "parallelism": f"dp{world}", "cuda_graph": use_graph,
                       "l2": "per-step working set (>10 GB of activations) far exceeds the 126 MB L2; no explicit flush"},
            "clocks": clocks, "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches_per_step) * args.steps, "gpu_launches_per_step": int(launches_per_step),
            "roofline": roof, "cpu_baseline": cpu, "inference": infer, "final_loss": round(final_loss, 4),
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
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=128, help="per-GPU batch")
    ap.add_argument("--seq", type=int, default=64)
    ap.add_argument("--fusion", default="basic", choices=["basic", "multiscale", "concat"],
                    help="fusion_type of the benchmarked model (default: the headline cross-attention block; "
                         "`multiscale` is what configs/ham_fusion_crossattn_v1.yml ships)")
    ap.add_argument("--ref-batch", type=int, default=8, help="samples per CPU step of the reference arm")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-inference", action="store_true", help="skip the eval-forward throughput measurement")
    ap.add_argument("--graph-multi-gpu", action="store_true", help="(default now; kept for old command lines)")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--dump-gemms", action="store_true", help="write per-shape GEMM timings to gpurun_out/gemm_shapes.txt")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

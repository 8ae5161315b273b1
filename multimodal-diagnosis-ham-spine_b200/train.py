"""Training step for the multimodal model on one B200 per process (the caller of the hot path:
scripts/train.py:349-394 in the reference): forward_features -> classifier -> loss -> backward -> gradient
all-reduce (data parallel) -> fused optimizer step over the flat parameter buffer.

* The whole step can be captured into ONE CUDA graph (`capture=True`): learning rate, step count and the
  dropout seed tick live in device memory, so replays stay correct.
* Data parallel = one process per GPU; gradients are averaged with NCCL over NVLink in two contiguous
  slices of the flat gradient buffer: [text encoder | fusion | head] is reduced on the communication
  stream while the ResNet trunk is still back-propagating, the image-encoder slice afterwards.
"""
import math

import torch
import torch.distributed as dist

from . import functional as Fm
from . import ops
from .parallel import GradSync, param_range, split_offset


def mibf_forward_loss(model, images, ids, mask, labels):
    """forward_loss for mibf_net.Resnet50WithOurs (mibf_net/train_resnet.py:21-41): batch dict -> three logit sets -> MP-Loss."""
    out = model({"transformed_image": images, "input_ids": ids, "attention_mask": mask})
    return model.cal_loss(out, labels), out["image_text"]


def connext_forward_loss(model, images, ids, mask, labels):
    """forward_loss for connext.OurClassfierConvnextV2 (ConNexT/models/pl_model_MOE2.py:144-154): logits -> cross entropy."""
    logits = model({"transformed_image": images, "input_ids": ids, "attention_mask": mask})
    return Fm.cross_entropy(logits, labels), logits


class Trainer:
    def __init__(self, model, optimizer="adamw", lr=2e-4, weight_decay=None, betas=(0.9, 0.999), eps=1e-8, momentum=0.0,
                 loss="ce", label_smoothing=0.02, focal_gamma=2.0, class_weights=None, process_group=None,
                 overlap_comm="backward", supcon_weight=0.0, supcon_stage="finetune", supcon_temperature=0.07, forward_loss=None,
                 comm_dtype="bf16", sm_reserve=None, bert_bucket_layers=0):
        """forward_loss: optional callable (model, images, ids, mask, labels) -> (loss, logits) for the model families whose
        call surface differs from MultimodalBaselineModel (MIBF-Net: batch dict + cal_loss, mibf_net/train_resnet.py:21-41;
        ConNexT: batch dict -> logits).  Everything else -- CUDA graph, gradient sync, fused optimizer -- is shared.
        Data parallel (world > 1): `comm_dtype` "bf16" (default) | "fp32" = dtype of the gradient buckets on the wire;
        `bert_bucket_layers` = BERT layers per early bucket (the text-encoder gradients leave in slices while its backward is
        still running; 0 = one slice after the whole text encoder); `sm_reserve` = SMs the persistent GEMM grids leave to the
        NCCL kernels while a bucket is in flight in "backward" mode (default: MDHS_SM_RESERVE or 0 -- measured on 2 x B200, a
        reserve of 16 costs more on the un-overlapped GEMMs than it gains on the overlapped ones).
        `momentum` only applies to optimizer="sgd": the default 0 is scripts/train.py:301-309's `optim.SGD(params, lr)`;
        MIBF-Net's recipe (mibf_net/train_resnet.py:136-139) passes momentum=0.9 explicitly."""
        self.model = model
        self.opt = optimizer.lower()
        if self.opt not in ("adamw", "adam", "sgd"):
            raise ValueError(f"Unsupported optimizer: {optimizer}")
        if weight_decay is None:
            weight_decay = 1e-2 if self.opt == "adamw" else 0.0   # torch defaults used by scripts/train.py:283-309
        self.lr, self.wd, self.betas, self.eps, self.momentum = lr, weight_decay, betas, eps, momentum
        self.loss_type, self.label_smoothing, self.focal_gamma = loss, label_smoothing, focal_gamma
        self.class_weights = class_weights
        # supervised contrastive term on the fused (B, hidden) features (scripts/train.py:341-346,363-383): stage "pretrain"
        # trains on SupCon alone, "finetune" adds supcon_weight * SupCon to the classification loss
        if supcon_stage not in ("pretrain", "finetune"):
            raise ValueError(f"Unsupported supcon stage: {supcon_stage}")
        self.supcon_weight, self.supcon_stage, self.supcon_temperature = float(supcon_weight), supcon_stage, supcon_temperature
        self.pg = process_group
        self.forward_loss = forward_loss
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        # the early buckets assume MultimodalBaselineModel's parameter order [image encoder | text encoder | fusion | head];
        # other model families reduce everything after backward
        # ... and assume ONE image-encoder / text-encoder backward per step: the global-local and multi-slice branches run the
        # trunk twice in train mode (model.py:292-315), so their stage hooks would fire before the gradients are final
        # (the gated model runs the whole feature path twice in train mode, model.py:257-271: same reason)
        self.gated = bool(getattr(model, "gate_enabled", False))
        twice = getattr(model, "global_local_enabled", False) or getattr(model, "sequence_enabled", False) or self.gated
        # overlap_comm: "backward" (default, or True) = buckets leave DURING backward ([text encoder | fusion | head] when the
        # BERT backward returns, ResNet layer4 / layer3 per stage, the rest at the end); "pipeline" = all buckets are reduced
        # after backward, the fused optimizer of bucket k alternating with the collective of bucket k+1; False / "none" = one
        # all-reduce, then the optimizer.  Measured (profiles/r02_summary.md section 5): 8 x B200 21.82 / 22.44 / 21.98 ms,
        # 2 x B200 22.03 / 21.99 / 22.29 ms against 20.65-21.4 ms on one GPU.  What the overlap costs: NCCL's CTAs need whole
        # SMs, and a statically scheduled persistent GEMM grid that loses even one SM to them runs a second round.
        mode = {True: "backward", False: "none", None: "none"}.get(overlap_comm, overlap_comm)
        if mode not in ("pipeline", "backward", "none"):
            raise ValueError(f"Unsupported overlap_comm: {overlap_comm}")
        self.comm_mode = mode if self.world > 1 else "none"
        self.overlap = self.comm_mode == "backward" and forward_loss is None and not twice
        import os as _os
        self.comm_dtype = {"bf16": torch.bfloat16, "fp32": None, "f32": None, None: None}[comm_dtype]
        self.sm_reserve = int(_os.environ.get("MDHS_SM_RESERVE", "0")) if sm_reserve is None else int(sm_reserve)
        self.bert_bucket_layers = int(bert_bucket_layers)
        # SMs the pipelined optimizer leaves to NCCL (its CTAs need an SM to themselves; pair with NCCL_MAX_CTAS <= comm_sms).
        # 0 (default) = normal optimizer grid: measured on 2 x B200, running the collective NEXT TO the HBM-bound optimizer
        # stretches the NCCL kernels 2.5x (320 vs 130 us per 54 MB bucket) and loses to letting them alternate
        # (21.81 vs 21.99 - 0.35 ms/step, profiles/r02_summary.md)
        self.comm_sms = int(_os.environ.get("MDHS_COMM_SMS", "0"))
        self.store = None
        self._graph = None
        self._static = None
        self._comm_works = []

    # ------------------------------------------------------------------ setup
    def _ensure(self, device):
        if self.store is not None:
            return
        st = self.model.store(device)
        self.store = st
        n = st.total
        self.m = torch.zeros(n, device=device)
        self.v = torch.zeros(n, device=device) if self.opt != "sgd" else None
        self.lr_dev = torch.full((1,), float(self.lr), device=device)
        self.step_dev = torch.zeros(1, device=device, dtype=torch.int32)
        # slice boundary for the overlapped all-reduce: everything from the text encoder on
        self.split = split_offset(st, getattr(self.model, "text_encoder", None))
        reserve = (lambda: ops.set_sm_reserve(self.sm_reserve)) if (self.overlap and self.sm_reserve > 0) else None
        self.sync = GradSync(st.grad, self.split if self.overlap else 0, self.pg, comm_dtype=self.comm_dtype,
                             on_first_reduce=reserve)
        # the optimizer only touches what torch.optim would: parameters with requires_grad (scripts/train.py builds the
        # optimizer from filter(requires_grad)) and, after the first backward, only those that actually received a gradient
        # (torch skips `grad is None`: the BERT pooler, dead q/k projections of the 1-token attention-pooling head, ...)
        self._opt_ranges = self._trainable_ranges(None)
        self._probed_unused = False
        self._buckets = self._pipeline_buckets(n)
        if self.world > 1:
            # same initial weights everywhere (rank 0 wins), like DistributedDataParallel's constructor
            dist.broadcast(st.flat, src=0, group=self.pg)
            st.refresh(force=True)

    def _trainable_ranges(self, skip_ids):
        """Merged [lo, hi) spans of the flat buffer that the optimizer updates.  Spans are merged across alignment padding
        only (padding elements are zero and stay zero); a frozen or unused parameter in between splits the span."""
        st = self.store
        spans, cur = [], None
        for p in st.params:                      # flat-buffer order
            o = st.offsets[id(p)]
            live = p.requires_grad and not (skip_ids and id(p) in skip_ids)
            if live:
                if cur is None:
                    cur = [o, o + p.numel()]
                else:
                    cur[1] = o + p.numel()
            elif cur is not None:
                spans.append(cur)
                cur = None
        if cur is not None:
            spans.append(cur)
        out = []
        for lo, hi in spans:                     # the kernels work on float4: round outwards inside the padded slot grid
            lo4, hi4 = lo // 4 * 4, min(st.total, (hi + 3) // 4 * 4)
            out.append((lo4, hi4))
        return out

    def _pipeline_buckets(self, total, sizes=(0.06, 0.12, 0.2, 0.2, 0.2, 0.22)):
        """[lo, hi) slices of the flat buffer for the pipelined reduce -> optimizer tail: a small first bucket (the optimizer
        starts early), boundaries on parameter slots (64-element aligned)."""
        st = self.store
        starts = sorted(st.offsets[id(p)] for p in st.params)
        cuts, acc = [0], 0.0
        for f in sizes[:-1]:
            acc += f
            target = acc * total
            c = min(starts, key=lambda o: abs(o - target))
            if c > cuts[-1]:
                cuts.append(c)
        cuts.append(total)
        return [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]

    def _opt_update(self, lo, hi, scale, g16):
        """Fused optimizer over the trainable spans inside [lo, hi)."""
        st = self.store
        bps = -self.comm_sms if (self.comm_mode == "pipeline" and self.comm_sms > 0) else 0
        for a, b in self._opt_ranges:
            a, b = max(a, lo), min(b, hi)
            if b <= a:
                continue
            if self.opt == "sgd":
                ops.sgd_flat(st.flat[a:b], st.grad[a:b], self.m[a:b] if self.momentum > 0 else None, st.shadow[a:b],
                             self.lr, self.momentum, self.wd, grad_scale=scale, lr_dev=self.lr_dev, step_dev=self.step_dev,
                             grads_bf16=None if g16 is None else g16[a:b], blocks_per_sm=bps)
            else:
                ops.adam_flat(st.flat[a:b], st.grad[a:b], self.m[a:b], self.v[a:b], st.shadow[a:b], self.lr,
                              self.betas[0], self.betas[1], self.eps, self.wd, 1, grad_scale=scale,
                              adamw=(self.opt == "adamw"), lr_dev=self.lr_dev, step_dev=self.step_dev,
                              grads_bf16=None if g16 is None else g16[a:b], blocks_per_sm=bps)

    def _probe_unused(self):
        """One-time (first eager step, after backward): parameters whose whole gradient is exactly zero never took part in
        the graph -- torch would have left `.grad = None` and its optimizers would skip them."""
        st = self.store
        flags = torch.stack([st.g32(p).count_nonzero() for p in st.params if p.requires_grad])
        live = [p for p in st.params if p.requires_grad]
        dead = {id(p) for p, n in zip(live, flags.tolist()) if n == 0}
        self._opt_ranges = self._trainable_ranges(dead)
        self._probed_unused = True

    def set_lr(self, lr):
        self.lr = lr
        if self.store is not None:
            self.lr_dev.fill_(float(lr))

    # ------------------------------------------------------------------ one step (eager)
    def _loss(self, logits, labels):
        if self.loss_type == "focal":
            return Fm.cross_entropy(logits, labels, self.class_weights, 0.0, True, self.focal_gamma)
        return Fm.cross_entropy(logits, labels, self.class_weights, self.label_smoothing, False, 2.0)

    def _step_impl(self, images, ids, mask, labels):
        st = self.store
        ops.step_begin(self.step_dev)
        hook = None
        img_eng = None
        if self.overlap and self.split > 0:
            eng = self.model.text_encoder._engine
            orig = eng.backward
            layers = self.model.text_encoder.model.encoder.layer
            nb = self.bert_bucket_layers

            def layer_done(li):
                # gradients of encoder layers [li, li + nb) are final: their slice leaves while the layers below still
                # back-propagate.  The first call also carries everything behind the encoder layers (fusion, head, pooler).
                if nb > 0 and li % nb == 0 and li > 0:
                    lo, _ = param_range(st, layers[li])
                    hi = self._bert_hi
                    self._bert_hi = lo
                    self.sync.reduce_range(lo, hi)

            main_stream = torch.cuda.current_stream()

            def bert_backward_then_reduce(ctx, dh, dtaps=None):
                # with the text encoder on its own stream (runtime.DUAL_STREAM) this runs in that stream's context: the
                # bucket below also carries the fusion / head gradients, whose kernels were enqueued on the step's stream
                cur = torch.cuda.current_stream()
                if cur != main_stream:
                    cur.wait_stream(main_stream)
                self._bert_hi = st.total
                eng.on_layer_backward_done = layer_done
                try:
                    orig(ctx, dh, dtaps)
                finally:
                    eng.on_layer_backward_done = None
                self.sync.reduce_range(self.split, self._bert_hi)   # embeddings + the lowest layers

            eng.backward = bert_backward_then_reduce
            hook = (eng, orig)
        if self.overlap:
            # ResNet stages finish in the order layer4 .. layer1: layer4 + layer3 hold 94 % of the trunk's parameters, so
            # their buckets leave while the (activation-heavy) early stages are still back-propagating
            img_eng = getattr(getattr(self.model, "image_encoder", None), "_engine", None)
            if img_eng is not None and hasattr(img_eng, "on_stage_backward_done"):
                net = img_eng.net
                stages = [net.layer1, net.layer2, net.layer3, net.layer4]

                def stage_done(li):
                    if li >= 2:
                        self.sync.reduce_range(*param_range(st, stages[li]))

                img_eng.on_stage_backward_done = stage_done
        try:
            if self.forward_loss is not None:
                loss, logits = self.forward_loss(self.model, images, ids, mask, labels)
            elif self.gated and self.supcon_weight == 0.0:
                # scripts/train.py:378 calls model(...): with the dual-expert gate that is NOT classifier(forward_features)
                logits = self.model(images, ids, mask)
                loss = self._loss(logits, labels)
            else:
                feats = self.model.forward_features(images, ids, mask)
                logits = self.model.classifier(feats)
                if self.supcon_weight > 0.0 and self.supcon_stage == "pretrain":
                    loss = Fm.supcon_loss(feats.float(), labels, self.supcon_temperature)
                else:
                    loss = self._loss(logits, labels)
                    if self.supcon_weight > 0.0:
                        loss = loss + self.supcon_weight * Fm.supcon_loss(feats.float(), labels, self.supcon_temperature)
            loss.backward()
        finally:
            if hook is not None:
                hook[0].backward = hook[1]
            if img_eng is not None:
                img_eng.on_stage_backward_done = None
        if not self._probed_unused and not torch.cuda.is_current_stream_capturing():
            self._probe_unused()            # on the LOCAL gradients, before any collective touches them
        if self.comm_mode == "pipeline":
            # every bucket goes on the wire in order; the optimizer of bucket k runs while bucket k + 1 is being reduced
            for lo, hi in self._buckets:
                self.sync.reduce_range(lo, hi)
            g16 = self.sync.reduced
            scale = 1.0 / self.world
            for i, (lo, hi) in enumerate(self._buckets):
                self.sync.wait(i)
                self._opt_update(lo, hi, scale, g16)
            self.sync.finish()
        else:
            scale = self.sync.finish()
            if self.sm_reserve > 0 and self.overlap:
                ops.set_sm_reserve(0)
            self._opt_update(0, st.total, scale, self.sync.reduced)
        st.mark_fresh()
        return loss.detach(), logits.detach()

    def step(self, images, ids, mask, labels):
        """Eager step on device tensors.  Returns (loss, logits) device tensors (no host sync)."""
        self._ensure(images.device)
        self.model.train()
        return self._step_impl(images, ids, mask, labels)

    # ------------------------------------------------------------------ CUDA-graph step
    def capture(self, images, ids, mask, labels, warmup=3):
        """Capture the whole step for fixed shapes.  Afterwards `replay(images, ids, mask, labels)` copies the new
        batch into the static input buffers and launches the graph.  NOTE: the `warmup` eager passes before the capture are
        REAL optimizer steps on the given batch (they also size the caching allocator and probe unused parameters)."""
        self._ensure(images.device)
        self.model.train()
        self._static = [t.clone() for t in (images, ids, mask, labels)]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._step_impl(*self._static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        # with NCCL in the step, other threads (the process-group watchdog) keep issuing CUDA calls while this thread
        # captures: only this thread's unsafe calls may invalidate the capture
        mode = "thread_local" if self.world > 1 else "global"
        with torch.cuda.graph(self._graph, capture_error_mode=mode):
            self._static_out = self._step_impl(*self._static)
        return self._static_out

    def release_graph(self):
        """Drop the captured graph (and its private memory pool); `step()` keeps working."""
        self._graph = None
        self._static_out = None

    def replay(self, images=None, ids=None, mask=None, labels=None):
        if images is not None:
            for dst, src in zip(self._static, (images, ids, mask, labels)):
                dst.copy_(src, non_blocking=True)
        self._graph.replay()
        return self._static_out


def warmup_cosine_lr(base_lr, epoch, warmup_epochs, total_epochs):
    """Learning rate of scripts/train.py:321-334's LambdaLR (linear warm-up then cosine decay).  The reference steps that
    scheduler once per BATCH with `warmup_steps = warmup_epochs * len(loader)`: pass step counts for all three arguments
    (`epoch` = global step, `warmup_epochs` = warm-up steps, `total_epochs` = total steps) to reproduce it exactly; with
    epoch counts it is the per-epoch variant."""
    if warmup_epochs > 0 and epoch < warmup_epochs:
        return base_lr * float(epoch + 1) / float(warmup_epochs)
    progress = (epoch - warmup_epochs) / max(1, total_epochs - warmup_epochs)
    return base_lr * 0.5 * (1.0 + math.cos(math.pi * progress))

"""Flat parameter storage for one model on one GPU.

All parameters of a model live in ONE fp32 buffer (master weights), with a same-shaped fp32 gradient
buffer and a bf16 shadow copy that the tensor-core kernels read.  `nn.Parameter.data` / `.grad` become
views into these buffers, so state_dict()/load_state_dict() and torch.optim keep working unchanged
(the state_dict key set is the drop-in contract), while
  * the optimizer step is one fused kernel over the flat buffer (csrc/optim.cu),
  * data-parallel gradient buckets are plain contiguous slices (parallel.py),
  * groups of parameters that one GEMM wants to see as a single operand (BERT q/k/v) are laid out
    back to back and exposed as one fused view without any copy.
Gradients are written by our backward kernels straight into the flat gradient buffer (they do not
travel through autograd's AccumulateGrad).
"""
import torch

from . import ops

_ALIGN = 64  # elements; keeps every view 256-byte aligned in fp32 and 128-byte aligned in bf16


class ParamStore:
    def __init__(self, module, device, groups=()):
        """groups: iterable of lists of parameters that must be contiguous (in that order)."""
        self.device = torch.device(device)
        params = []
        seen = set()
        order = []
        grouped = {}
        for grp in groups:
            for p in grp:
                grouped[id(p)] = grp
        for p in module.parameters():
            if id(p) in seen:
                continue
            if id(p) in grouped:
                for q in grouped[id(p)]:
                    if id(q) not in seen:
                        seen.add(id(q))
                        order.append((q, False))
                # the next group member must directly follow: mark packing (no alignment gap)
            else:
                seen.add(id(p))
                order.append((p, True))
        # members of a group after the first are packed without alignment padding
        packed = set()
        for grp in groups:
            for q in grp[1:]:
                packed.add(id(q))
        offsets = {}
        off = 0
        for p, _ in order:
            if id(p) not in packed:
                off = (off + _ALIGN - 1) // _ALIGN * _ALIGN
            offsets[id(p)] = off
            off += p.numel()
            params.append(p)
        self.total = (off + _ALIGN - 1) // _ALIGN * _ALIGN
        self.params = params
        self.offsets = offsets
        self.flat = torch.zeros(self.total, device=self.device, dtype=torch.float32)
        self.grad = torch.zeros(self.total, device=self.device, dtype=torch.float32)
        self.shadow = torch.zeros(self.total, device=self.device, dtype=torch.bfloat16)
        with torch.no_grad():
            for p in params:
                o = offsets[id(p)]
                view = self.flat[o:o + p.numel()].view(p.shape)
                view.copy_(p.data.to(self.device, torch.float32))
                p.data = view
                p.grad = self.grad[o:o + p.numel()].view(p.shape)
        self._versions = None
        self._packers = []  # callables run after every shadow refresh (conv weight re-layout)
        # dummy differentiable leaf: makes autograd call the backward of Functions whose real
        # parameters do not pass through autograd
        self.anchor = torch.zeros(1, device=self.device, dtype=torch.float32, requires_grad=True)
        self.refresh(force=True)

    # ------------------------------------------------------------------ views
    def owns(self, p):
        o = self.offsets.get(id(p))
        return o is not None and p.data.data_ptr() == self.flat.data_ptr() + 4 * o

    def valid(self):
        return all(self.owns(p) for p in self.params)

    def w16(self, p):
        o = self.offsets[id(p)]
        return self.shadow[o:o + p.numel()].view(p.shape)

    def g32(self, p):
        o = self.offsets[id(p)]
        return self.grad[o:o + p.numel()].view(p.shape)

    def fused(self, plist, shape):
        """One view over parameters declared as a contiguous group: (master fp32, shadow bf16, grad fp32)."""
        o = self.offsets[id(plist[0])]
        n = sum(p.numel() for p in plist)
        oo = o
        for p in plist:
            assert self.offsets[id(p)] == oo, "parameters are not contiguous in the flat buffer"
            oo += p.numel()
        return (self.flat[o:o + n].view(shape), self.shadow[o:o + n].view(shape), self.grad[o:o + n].view(shape))

    # ------------------------------------------------------------------ freshness of the bf16 copies
    def add_packer(self, fn):
        self._packers.append(fn)
        fn()

    def mark_fresh(self):
        """Called by the fused optimizer, which rewrites the shadow itself."""
        for fn in self._packers:
            fn()
        self._versions = [p._version for p in self.params]

    def refresh(self, force=False):
        """Recast master -> bf16 when any parameter changed through torch (load_state_dict, torch.optim)."""
        versions = [p._version for p in self.params]
        if force or versions != self._versions:
            ops.cast_f32_bf16(self.flat, out=self.shadow)
            for fn in self._packers:
                fn()
            self._versions = versions

    def attach_grads(self):
        """Stock-loop support (scripts/train.py:364-387: `optimizer.zero_grad()` ... `loss.backward()` ... `optimizer.step()`):
        torch's `zero_grad(set_to_none=True)` drops `.grad`, while our backward kernels accumulate into the flat gradient
        buffer.  Called at the start of every grad-enabled forward: parameters whose `.grad` was dropped get their slice of
        the flat buffer ZEROED (the caller asked for fresh gradients) and re-attached, so torch optimizers see them again.
        With the fused Trainer nothing is ever detached and this is a no-op scan."""
        missing = [p for p in self.params if p.requires_grad and p.grad is None]
        if not missing:
            return
        n_train = sum(1 for p in self.params if p.requires_grad)
        if len(missing) == n_train:
            self.grad.zero_()
        with torch.no_grad():
            for p in missing:
                o = self.offsets[id(p)]
                g = self.grad[o:o + p.numel()].view(p.shape)
                if len(missing) != n_train:
                    g.zero_()
                p.grad = g

    def zero_grad(self):
        self.grad.zero_()


# ------------------------------------------------------------------ side stream for independent backward work
# Weight-gradient GEMMs do not feed the data-gradient chain: running them on a second stream could let their CTAs start on
# the SMs that the persistent dgrad kernel leaves idle in its last, partially filled wave.  fork() / join() also work inside
# CUDA-graph capture.  MEASURED (round 1, B200, config 2): no gain -- 22.89 ms/step with the overlap vs 22.76 without -- so it
# is OFF by default (MDHS_OVERLAP_WGRAD=1 enables it for experiments).
import os as _os

_SIDE = {}
OVERLAP_WGRAD = _os.environ.get("MDHS_OVERLAP_WGRAD", "0") == "1"


# The image and the text encoder do not depend on each other: the text encoder (forward, and through autograd's stream
# bookkeeping also its backward) can run on a second stream, so that its tensor-bound GEMMs and small attention grids overlap
# the HBM-bound BatchNorm passes of the ResNet trunk.  Needs the GEMM's dynamic work distribution (a statically scheduled
# persistent grid that shares SMs with another stream's kernels runs a second round: 20.3 -> 26.9 ms/step, measured; with the
# work counter: 20.6 -> 19.2 ms/step).  MDHS_DUAL_STREAM=0 keeps everything on one stream.
DUAL_STREAM = _os.environ.get("MDHS_DUAL_STREAM", "1") == "1"
_BRANCH = {}


def fork_branch():
    """Second branch stream that has waited for everything enqueued so far on the current stream."""
    dev = torch.cuda.current_device()
    s = _BRANCH.get(dev)
    if s is None:
        s = _BRANCH[dev] = torch.cuda.Stream(device=dev)
    s.wait_stream(torch.cuda.current_stream())
    return s


def record_on_current(obj):
    """Tensors produced on another stream and consumed on the current one: tell the caching allocator."""
    cur = torch.cuda.current_stream()
    if torch.is_tensor(obj):
        if obj.is_cuda:
            obj.record_stream(cur)
    elif isinstance(obj, dict):
        for v in obj.values():
            record_on_current(v)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            record_on_current(v)


def join_after_backward(main, branch):
    """Called from a backward that runs on `branch`: when the whole backward pass has been enqueued, `main` (the stream the
    step runs on: optimizer, gradient collectives, the user's next kernels) waits for `branch`.  Parameter gradients are
    written by our kernels, not by AccumulateGrad nodes, so autograd's own end-of-backward stream sync does not cover them."""
    from torch.autograd import Variable
    Variable._execution_engine.queue_callback(lambda: main.wait_stream(branch))


class BranchGate(torch.autograd.Function):
    """Identity on a tensor that leaves the branch stream: its backward is the first node of the branch's backward and
    registers the end-of-backward join."""

    @staticmethod
    def forward(ctx, x, main, branch):
        ctx.streams = (main, branch)
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        join_after_backward(*ctx.streams)
        return g, None, None


def gate_branch_outputs(obj, main, branch):
    if torch.is_tensor(obj):
        return BranchGate.apply(obj, main, branch) if obj.requires_grad else obj
    if isinstance(obj, dict):
        return {k: gate_branch_outputs(v, main, branch) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(gate_branch_outputs(v, main, branch) for v in obj)
    return obj


def fork_side():
    """Side stream that has waited for everything enqueued so far on the current stream."""
    dev = torch.cuda.current_device()
    s = _SIDE.get(dev)
    if s is None:
        s = _SIDE[dev] = torch.cuda.Stream(device=dev)
    s.wait_stream(torch.cuda.current_stream())
    return s


def join_side(s):
    torch.cuda.current_stream().wait_stream(s)

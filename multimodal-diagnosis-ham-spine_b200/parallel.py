"""Data-parallel gradient synchronisation over the flat gradient buffer (one process per GPU).

The reference's only multi-GPU path is DistributedDataParallel over NCCL (mibf_net/train_resnet.py:84-88,
133-134).  Here gradients already live in ONE contiguous fp32 buffer (runtime.ParamStore), so a bucket is a
slice: the tail slice [split, total) -- text encoder, fusion, head, whose gradients are complete as soon as
the BERT backward returns -- is all-reduced asynchronously while the ResNet trunk is still back-propagating;
the head slice [0, split) follows when backward ends.  The sum is turned into a mean by the optimizer
kernel (grad_scale = 1/world), so no extra pass touches the buffer.

`comm_dtype=torch.bfloat16` (the Trainer's default at world > 1): every bucket is cast into a bf16 communication
buffer by one kernel right before its all-reduce -- half the bytes on the wire and half the time the collective's CTAs share
the GPU with the backward kernels -- and the fused optimizer reads the reduced gradient straight from that buffer.

Backend-agnostic on purpose: NCCL over NVLink on the GPUs, gloo in the CPU unit tests.
"""
import torch
import torch.distributed as dist


def split_offset(store, tail_module):
    """First flat-buffer offset owned by `tail_module` (everything from there on forms the early bucket)."""
    if tail_module is None:
        return 0
    offs = [store.offsets[id(p)] for p in tail_module.parameters() if id(p) in store.offsets]
    return min(offs) if offs else 0


class GradSync:
    """Buckets are arbitrary [lo, hi) slices of the flat gradient buffer, reduced asynchronously as soon as the caller
    declares them final (`reduce_range`); `finish()` reduces whatever is left and waits for everything."""

    def __init__(self, flat_grad, split, group=None, comm_dtype=None, on_first_reduce=None):
        self.grad = flat_grad
        self.total = flat_grad.numel()
        self.split = max(0, min(int(split), self.total))
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self._works = []
        self._done = []      # [lo, hi) ranges already handed to the collective this step
        # bf16 buckets: `comm` holds the (cast, then all-reduced) gradient; None = reduce the fp32 buffer in place
        self.comm = None
        if comm_dtype is not None and comm_dtype != flat_grad.dtype and self.world > 1:
            self.comm = torch.zeros(self.total, device=flat_grad.device, dtype=comm_dtype)
        self.on_first_reduce = on_first_reduce   # called once per step right before the first collective is enqueued
        self._started = False

    @property
    def reduced(self):
        """Buffer that holds the all-reduced gradient after finish(): the bf16 comm buffer, or None (= the fp32 grad)."""
        return self.comm

    def _cast(self, lo, hi):
        if self.grad.is_cuda:
            from . import ops
            ops.cast_f32_bf16(self.grad[lo:hi], out=self.comm[lo:hi])
        else:                       # gloo unit tests on CPU tensors
            self.comm[lo:hi].copy_(self.grad[lo:hi])

    def _reduce(self, lo, hi):
        if self.world > 1 and hi > lo:
            if not self._started:
                self._started = True
                if self.on_first_reduce is not None:
                    self.on_first_reduce()
            buf = self.grad
            if self.comm is not None:
                self._cast(lo, hi)
                buf = self.comm
            self._works.append(dist.all_reduce(buf[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def reduce_range(self, lo, hi):
        """Call when every gradient in [lo, hi) is final.  Ranges must not overlap within one step."""
        lo, hi = max(0, int(lo)), min(self.total, int(hi))
        if hi <= lo:
            return
        for a, b in self._done:
            assert hi <= a or lo >= b, "overlapping gradient buckets"
        self._done.append((lo, hi))
        self._reduce(lo, hi)

    def reduce_tail(self):
        """Call when every gradient in [split, total) is final (after the text-encoder backward)."""
        self.reduce_range(self.split, self.total)

    def wait(self, i):
        """Block the current stream until the i-th collective issued this step has completed (pipelined optimizer tail)."""
        if i < len(self._works):
            self._works[i].wait()

    def finish(self):
        """Call after backward: reduces what is left and waits for all outstanding collectives."""
        pos = 0
        for a, b in sorted(self._done):
            self._reduce(pos, a)
            pos = b
        self._reduce(pos, self.total)
        for w in self._works:
            w.wait()
        self._works = []
        self._done = []
        self._started = False
        return 1.0 / self.world   # scale that turns the summed gradient into the data-parallel mean


def param_range(store, module):
    """[lo, hi) of the flat buffer covered by `module`'s parameters (contiguous by construction of ParamStore)."""
    spans = [(store.offsets[id(p)], store.offsets[id(p)] + p.numel()) for p in module.parameters() if id(p) in store.offsets]
    if not spans:
        return 0, 0
    return min(a for a, _ in spans), max(b for _, b in spans)

"""Data-parallel gradient synchronisation over the flat gradient buffer (one process per GPU).

The reference's only multi-GPU path is DistributedDataParallel over NCCL (mibf_net/train_resnet.py:84-88,
133-134).  Here gradients already live in ONE contiguous fp32 buffer (runtime.ParamStore), so a bucket is a
slice: the tail slice [split, total) -- text encoder, fusion, head, whose gradients are complete as soon as
the BERT backward returns -- is all-reduced asynchronously while the ResNet trunk is still back-propagating;
the head slice [0, split) follows when backward ends.  The sum is turned into a mean by the optimizer
kernel (grad_scale = 1/world), so no extra pass touches the buffer.

Backend-agnostic on purpose: NCCL over NVLink on the GPUs, gloo in the CPU unit tests.
"""
import torch.distributed as dist


def split_offset(store, tail_module):
    """First flat-buffer offset owned by `tail_module` (everything from there on forms the early bucket)."""
    if tail_module is None:
        return 0
    offs = [store.offsets[id(p)] for p in tail_module.parameters() if id(p) in store.offsets]
    return min(offs) if offs else 0


class GradSync:
    def __init__(self, flat_grad, split, group=None):
        self.grad = flat_grad
        self.total = flat_grad.numel()
        self.split = max(0, min(int(split), self.total))
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self._works = []

    def _reduce(self, lo, hi):
        if self.world > 1 and hi > lo:
            self._works.append(dist.all_reduce(self.grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def reduce_tail(self):
        """Call when every gradient in [split, total) is final (after the text-encoder backward)."""
        self._tail_done = True
        self._reduce(self.split, self.total)

    def finish(self):
        """Call after backward: reduces what is left and waits for all outstanding collectives."""
        if getattr(self, "_tail_done", False):
            self._reduce(0, self.split)
        else:
            self._reduce(0, self.total)
        for w in self._works:
            w.wait()
        self._works = []
        self._tail_done = False
        return 1.0 / self.world   # scale that turns the summed gradient into the data-parallel mean

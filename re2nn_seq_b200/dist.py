"""Data parallelism over independent sequences.

The reference is single-process; its only multi-GPU story is one job per GPU (tools/builder.py:28-60).
Sequences are independent units of the hot path, so the batch is sharded across ranks with NO collective
on the data path.  Training adds exactly one collective per step: an all-reduce(SUM) of a flat fp32
bucket holding every trainable gradient (NCCL over NVLink/NVSwitch; gloo in the CPU tests).

Loss semantics under sharding (must match the single-process reference):
  * CRF loss is a SUM over sequences (baselines/crf.py:99,250,260)  -> gradients simply add.
  * CE loss is a MEAN over all valid tokens of the global batch (nn.CrossEntropyLoss default,
    farnn/model_decompose.py:80) -> every rank divides by the GLOBAL token count (`global_tokens`), then
    gradients add.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, world, rank):
    """Contiguous, balanced slice [lo, hi) of n sequences for `rank` (first n % world ranks get one extra)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(tensors, world, rank):
    """Slice every batch-major tensor of a global batch to this rank's sequences."""
    n = tensors[0].shape[0]
    lo, hi = shard_bounds(n, world, rank)
    return [t[lo:hi] for t in tensors]


def global_token_count(lengths, group=None):
    """Total valid tokens over all ranks (one scalar all-reduce)."""
    n = lengths.sum().to(torch.int64).reshape(1).clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(n, op=dist.ReduceOp.SUM, group=group)
    return int(n.item())


class GradBucket:
    """Flat fp32 bucket over the trainable parameters of a module; one all-reduce(SUM) per step.

    Every parameter's ``.grad`` IS a view of its slice of the bucket (set once, kept across steps), so the backward
    pass accumulates straight into the bucket and the collective runs on it in place: no pack / unpack copies.
    ``zero_grad()`` is one memset of the bucket (use it instead of setting ``p.grad = None``)."""

    def __init__(self, module, group=None):
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.group = group
        self.sizes = [p.numel() for p in self.params]
        self.total = sum(self.sizes)
        dev = self.params[0].device if self.params else torch.device('cpu')
        self.flat = torch.zeros((self.total,), dtype=torch.float32, device=dev)
        self.views = []
        off = 0
        for p, n in zip(self.params, self.sizes):
            self.views.append(self.flat[off:off + n].view_as(p))
            off += n
        self._bind()

    def _bind(self):
        for p, v in zip(self.params, self.views):
            if p.grad is not v:
                if p.grad is not None:          # a gradient produced before the bucket existed (or after grad = None)
                    v.copy_(p.grad)
                p.grad = v

    def zero_grad(self):
        self.flat.zero_()
        self._bind()

    def all_reduce(self):
        """all_reduce(SUM) of the bucket in place.  Returns the number of bytes reduced."""
        self._bind()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        return self.total * 4


def all_reduce_loss(loss, group=None):
    """Sum of the per-rank loss contributions (CRF: per-shard sums; CE: per-shard sums / global_tokens)."""
    out = loss.detach().clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out

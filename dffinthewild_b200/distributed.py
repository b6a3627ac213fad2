"""Data-parallel training plumbing for the depth-from-focus path: one process per GPU, one flat gradient bucket, ONE
all-reduce per step (NCCL over NVLink 5 / NVSwitch on the GPU box, gloo in the CPU tests).

What it replaces: `nn.DataParallel`'s per-step `broadcast_coalesced` of 16.18 MB of parameters and `reduce_add_coalesced` of
16.16 MB of gradients to GPU 0 (reference call sites train_codes/train_code_Defocus.py:63,158,167).
Each rank owns a replica, so the broadcast disappears; the gradient reduction becomes a single all-reduce of one contiguous
fp32 buffer (4,016,592 elements: the 12 parameters / 22,240 elements that never receive a gradient, SURVEY.md §8a row 13,
stay out of it).

BatchNorm statistics stay per rank, exactly as DataParallel keeps them per replica (no SyncBN is introduced).

Loss normalisation: the reference takes the masked-MSE mean over the valid pixels of the *gathered global* batch
(train_code_Defocus.py:17-19,160-164).  A rank whose loss is the mean over its own n_r valid pixels must weight its gradient
by n_r / sum(n): `allreduce_gradients(weight=n_r)` scales the bucket by n_r, all-reduces the bucket together with the weight,
and divides by the summed weight - one collective, identical to the global-batch mean.
"""
import torch
import torch.distributed as dist


class GradBucket:
    """Flat fp32 gradient buffer; every used parameter's `.grad` is a view into it."""

    def __init__(self, module, skip=()):
        self.params = [p for n, p in module.named_parameters() if p.requires_grad and n not in set(skip)]
        if not self.params:
            raise ValueError("GradBucket: no parameters")
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n + 1, dtype=torch.float32, device=dev)   # last element carries the loss weight
        o = 0
        for p in self.params:
            p.grad = self.flat[o:o + p.numel()].view_as(p)
            o += p.numel()
        self.numel = n

    def zero(self):
        """`optimizer.zero_grad(set_to_none=False)` equivalent that keeps the views alive."""
        self.flat.zero_()

    def rebind(self):
        """Make every `.grad` a view of the flat buffer again.  `optimizer.zero_grad()` / `model.zero_grad()` default to
        `set_to_none=True`, which drops the views: a gradient autograd then allocated on its own is copied into its slot and
        re-bound; a parameter without a gradient contributes zeros."""
        o = 0
        for p in self.params:
            n = p.numel()
            view = self.flat[o:o + n].view_as(p)
            g = p.grad
            if g is None:
                view.zero_()
                p.grad = view
            elif g.data_ptr() != view.data_ptr() or g.dtype != torch.float32:
                if g.shape != p.shape:
                    raise ValueError("GradBucket: foreign gradient of shape %s for a parameter of shape %s" % (tuple(g.shape), tuple(p.shape)))
                view.copy_(g)
                p.grad = view
            o += n

    def allreduce_gradients(self, weight=1.0, group=None):
        """Weighted average of the gradients over all ranks, in place, with ONE all-reduce."""
        self.rebind()
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        self.flat[:self.numel].mul_(float(weight))
        self.flat[self.numel] = float(weight)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        # (no rank has a valid pixel: the summed weight is 0 and so is every gradient — keep the zeros instead of 0/0)
        self.flat[:self.numel].div_(self.flat[self.numel].clamp_min(torch.finfo(torch.float32).tiny))


def unused_parameter_names(module):
    """The 12 parameters the forward never touches (`redir3`, `pre_conv`; reference :244, :285-286)."""
    return [n for n, _ in module.named_parameters() if ".redir3." in n or ".pre_conv." in n]


def shard_batch(n_items, rank, world):
    """Indices of the focal stacks rank `rank` owns: {rank, rank+world, ...}."""
    return list(range(rank, n_items, world))

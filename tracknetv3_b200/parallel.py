"""Data-parallel training across the GPUs of one box: one process per GPU, replicas hold identical
parameters, the only collective is ONE allreduce of the flat fp32 gradient bucket (45.36 MB for
TrackNet(27, 8)) over NCCL/NVLink per step. The reference has no multi-GPU code;
semantics follow torch DDP's defaults: gradients averaged over ranks, BatchNorm statistics per replica.
"""
import os

import torch
import torch.distributed as dist
from torch._utils import _flatten_dense_tensors, _unflatten_dense_tensors


def broadcast_module(module, src=0):
    """Make every rank's parameters and buffers equal to rank ``src``'s (what DDP does at construction)."""
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src)


class _GradSplit:
    """What TrackNet's backward and the bucket share when the allreduce overlaps the backward: the event recorded between
    the two ranges of the backward pass and the index of the first parameter whose gradient is final at that event."""

    def __init__(self):
        self.event = torch.cuda.Event()
        self.first_param = None


class GradBucket:
    """Flat gradient bucket: ``allreduce()`` averages all parameter gradients across ranks - in one call, or
    (``overlap=True``, NCCL, a module that hands its gradients out as one block and runs its backward in two ranges:
    TrackNet) in two: the bottleneck / decoder / predictor part, 85 % of the bytes, starts on a side stream as soon as the
    first range of the backward has produced it and runs under the encoder's backward; the small encoder part follows the
    backward on the caller's stream."""

    def __init__(self, module, process_group=None, overlap=False):
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.split = None
        if os.environ.get("TNB_ALLREDUCE_OVERLAP", "1") == "0":  # ablation: one allreduce after the backward
            overlap = False
        if (overlap and self.world > 1 and hasattr(module, "_grad_split") and dist.get_backend(process_group) == "nccl"
                and len(self.params) == len(list(module.parameters()))):
            self.split = module._grad_split = _GradSplit()
            self.side = torch.cuda.Stream()

    @staticmethod
    def _shared_flat(grads):
        """The gradients as ONE tensor without copying, if they already are back-to-back slices of one allocation in
        parameter order (TrackNet's backward hands them out that way), else None."""
        if not grads or any(not g.is_contiguous() for g in grads):
            return None
        st = grads[0].untyped_storage()
        off = grads[0].storage_offset()
        first = off
        for g in grads:
            if g.dtype != grads[0].dtype or g.untyped_storage().data_ptr() != st.data_ptr() or g.storage_offset() != off:
                return None
            off += g.numel()
        return torch.empty(0, dtype=grads[0].dtype, device=grads[0].device).set_(st, first, (off - first,))

    def allreduce(self):
        if self.world == 1:
            return
        grads = [p.grad for p in self.params if p.grad is not None]
        flat = self._shared_flat(grads)
        if (flat is not None and self.split is not None and self.split.first_param is not None
                and len(grads) == len(self.params)):
            off = sum(g.numel() for g in grads[:self.split.first_param])
            head, tail = flat[:off], flat[off:]
            self.side.wait_event(self.split.event)           # the tail is final once the backward's first range is done
            with torch.cuda.stream(self.side):
                w_tail = dist.all_reduce(tail, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
            w_head = dist.all_reduce(head, op=dist.ReduceOp.AVG, group=self.group, async_op=True)  # after the whole backward
            w_tail.wait()                                    # the caller's stream continues after both
            w_head.wait()
            self.split.first_param = None
            return
        if flat is not None:  # in place: no flatten / copy-back passes around the collective
            if dist.get_backend(self.group) == "nccl":  # ncclAvg: the division rides inside the collective
                dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
            else:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
                flat.div_(self.world)
            return
        flat = _flatten_dense_tensors(grads)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat.div_(self.world)
        for g, f in zip(grads, _unflatten_dense_tensors(flat, grads)):
            g.copy_(f)


def shard_batch(n_total, rank, world):
    """Contiguous sample range of a global batch owned by ``rank`` (pure data parallelism)."""
    per = n_total // world
    rem = n_total % world
    start = rank * per + min(rank, rem)
    return start, start + per + (1 if rank < rem else 0)

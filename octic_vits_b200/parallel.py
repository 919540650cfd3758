"""Batch-sharded data parallelism for the octic ViT step (reference: DDP in deit/main.py:354-359).

Every op of the path is per image, so ranks own disjoint slices of the batch and a full replica of the parameters;
the only exchange per step is the gradient all-reduce.  Gradients live in ONE flat fp32 buffer (the `.grad` of every
trainable parameter is a view into it) so the exchange is a single NCCL all-reduce over NVLink/NVSwitch -- no
per-parameter bookkeeping and no bucket scheduling.  Frozen parameters (cls_token.1-4, reference model.py:99-106) are
simply not part of the buffer, which is what the reference needs find_unused_parameters=True for.
"""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def shard_batch(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, stop) of the images owned by `rank` (contiguous, sizes differ by at most one)."""
    base, rem = divmod(global_batch, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class FlatGrads:
    """Flat fp32 gradient buffer + all-reduce.  Usage:
        fg = FlatGrads(model.parameters())
        fg.zero(); loss.backward(); fg.all_reduce()      # grads are now the mean over ranks
    """

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self) -> None:
        self.flat.zero_()

    def all_reduce(self, average: bool = True) -> None:
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat)
            if average:
                self.flat.div_(dist.get_world_size())

    def nbytes(self) -> int:
        return self.flat.numel() * 4

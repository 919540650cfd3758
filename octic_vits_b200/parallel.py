"""Batch-sharded data parallelism for the octic ViT step (reference: DDP in deit/main.py:354-359).

Every op of the path is per image, so ranks own disjoint slices of the batch and a full replica of the parameters;
the only exchange per step is the gradient all-reduce.  Gradients live in ONE flat fp32 buffer (the `.grad` of every
trainable parameter is a view into it) so the exchange is a single NCCL all-reduce over NVLink/NVSwitch -- no
per-parameter bookkeeping and no bucket scheduling.  Frozen parameters (cls_token.1-4, reference model.py:99-106) are
simply not part of the buffer, which is what the reference needs find_unused_parameters=True for.
"""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional, Tuple

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

    def __init__(self, params: Iterable[torch.nn.Parameter], fuse_accumulation: bool = True):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        # every view starts on a 16-byte boundary (the fused optimizer streams the buffer with 16-byte accesses)
        self.offsets: List[int] = []
        off = 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + 3) // 4 * 4
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        for p, o in zip(self.params, self.offsets):
            p.grad = self.flat[o:o + p.numel()].view_as(p)
        if fuse_accumulation:
            # the wgrad kernels add straight into these views (no zero fill + autograd accumulation kernel per weight)
            from . import functional as OF
            OF.ACCUMULATE_INTO_GRAD = True

    def zero(self) -> None:
        self.flat.zero_()

    def all_reduce(self, average: bool = True) -> None:
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            if average and dist.get_backend() == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)      # the division happens inside the NCCL kernel
            else:
                dist.all_reduce(self.flat)
                if average:
                    self.flat.div_(dist.get_world_size())

    def nbytes(self) -> int:
        return self.flat.numel() * 4


class GraphedTrainStep:
    """One training step -- zero grads, forward, loss, backward -- captured ONCE in a CUDA graph and replayed
    (whole-step orchestration).  The eager step of the headline model issues ~2000 kernel
    launches from Python; at ~150 ms per step the launch thread is the bottleneck in the backward of the small
    kernels, and the graph removes that (and every per-call tensor-map encode) from the critical path.

        fg = FlatGrads(model.parameters())
        step = GraphedTrainStep(model, fg, images.shape)         # warm-up + capture
        loss = step(images, targets)                             # copies into the static inputs, replays, all-reduces

    Inputs may live on the host (pinned) or on the device; gradients land in `fg.flat` (the parameters' .grad views).
    Input pipeline: `stage(images, targets)` starts the host -> device copy of a batch on a copy stream (it overlaps the
    step that is still running) and `run()` consumes the staged batch; `step(images, targets)` = stage + run.
    The gradient all-reduce stays outside the graph: it is a single NCCL call per step, followed -- when an
    `optimizer` (optim.FusedOptimizer) is given -- by its 2-4 launches; the weight packs are then part of the graph so
    that every replay sees the updated parameters.  If capture is impossible
    (e.g. an op that synchronises), `graphed` is False and every call runs the eager step instead.
    """

    def __init__(self, model: torch.nn.Module, flat_grads: FlatGrads, image_shape, *, target_dtype=torch.long,
                 loss_fn: Optional[Callable] = None, warmup: int = 3, use_graph: bool = True, optimizer=None):
        self.model, self.fg, self.optimizer = model, flat_grads, optimizer
        self.loss_fn = loss_fn or torch.nn.functional.cross_entropy
        dev = flat_grads.flat.device
        self.img = torch.zeros(tuple(image_shape), device=dev)
        self.tgt = torch.zeros(image_shape[0], dtype=target_dtype, device=dev)
        self._stage_img, self._stage_tgt = torch.zeros_like(self.img), torch.zeros_like(self.tgt)
        self._copy_stream = torch.cuda.Stream(device=dev)
        self._staged = torch.cuda.Event()
        self._consumed = torch.cuda.Event()
        self._consumed.record()
        self.graph, self.loss, self.graphed = None, None, False
        if not use_graph:
            return
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                     # packs, attribute settings and allocator warm-up happen here
            for _ in range(max(1, warmup)):
                self._eager()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        if optimizer is not None:
            # parameters change between replays: capture the fp32 -> bf16 weight packs inside the graph (cache miss
            # on first use of every weight) so that each replay re-packs from the current values
            from . import functional as OF
            OF.bump_param_epoch()
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                loss = self._eager()
            self.graph, self.loss, self.graphed = graph, loss, True
        except Exception as e:                            # noqa: BLE001 - any capture failure -> eager path, reported
            self.capture_error = repr(e)
            torch.cuda.synchronize(dev)

    def _eager(self) -> torch.Tensor:
        from . import ops
        ops.begin_step()                  # one zero fill for all red.add scratch of the step
        self.fg.flat.zero_()
        loss = self.loss_fn(self.model(self.img), self.tgt)
        loss.backward()
        return loss

    def stage(self, images: torch.Tensor, targets: torch.Tensor) -> None:
        """Start copying a batch into the staging buffers on the copy stream (after the previous batch was consumed)."""
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._consumed)
            self._stage_img.copy_(images, non_blocking=True)
            self._stage_tgt.copy_(targets, non_blocking=True)
            self._staged.record()

    def run(self) -> torch.Tensor:
        """One step on the staged batch (device -> device copy into the graph's static inputs, replay, all-reduce)."""
        cur = torch.cuda.current_stream(self.img.device)
        cur.wait_event(self._staged)
        self.img.copy_(self._stage_img, non_blocking=True)
        self.tgt.copy_(self._stage_tgt, non_blocking=True)
        self._consumed.record(cur)
        return self._launch()

    def __call__(self, images: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        self.img.copy_(images, non_blocking=True)
        self.tgt.copy_(targets, non_blocking=True)
        return self._launch()

    def _launch(self) -> torch.Tensor:
        if self.graphed:
            self.graph.replay()
            loss = self.loss
        else:
            loss = self._eager()
        self.fg.all_reduce()
        if self.optimizer is not None:
            self.optimizer.step()             # consumes the all-reduced flat gradients (optim.FusedOptimizer)
        return loss

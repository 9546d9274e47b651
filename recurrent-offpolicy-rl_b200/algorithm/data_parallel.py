"""Trajectory-sharded data parallelism for the update.

Every rank holds full replicas and updates on its own rows of the batch; what crosses ranks is
  * the gradients: SUM all-reduce of each model's flat gradient arena, cut into contiguous buckets that are
    reduced on a side stream as soon as the backward pass has produced every gradient inside them, so the
    transfer overlaps the rest of the backward (the heads' gradients are ready while the encoder is still
    back-propagating);
  * `n_valid` (SUM) and the Q-value guard bounds (MIN / MAX) before the backward, because the losses are
    normalised by the GLOBAL valid-step count (ref: offpolicy_rnn/algorithm/sac_full_length_rnn_ensembleQ.py:80-81,
    391) and the guard state must be identical on all ranks (ref: offpolicy_rnn/utility/q_value_guard.py:22-38).
Gradients are summed, not averaged: each rank's loss already carries the 1 / n_valid_global factor.

The module is device-agnostic (NCCL on GPUs, gloo in the CPU tests); it contains no arithmetic of the update.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_rows(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row block [lo, hi) of rank `rank`; the first n_rows % world ranks take one extra row."""
    base, extra = divmod(n_rows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def plan_buckets(offsets: Sequence[int], sizes: Sequence[int], total: int, bucket_elems: int) -> List[Tuple[int, int, List[int]]]:
    """Cut a flat arena into contiguous buckets of about `bucket_elems` elements on parameter boundaries.
    Returns [(start, end, [parameter indices])]; the buckets tile [0, total) exactly."""
    buckets, start, members = [], 0, []
    for i, (off, n) in enumerate(zip(offsets, sizes)):
        members.append(i)
        end = offsets[i + 1] if i + 1 < len(offsets) else total
        if end - start >= bucket_elems or i + 1 == len(offsets):
            buckets.append((start, end, members))
            start, members = end, []
    return buckets


class BucketedGradSync:
    """All-reduce (SUM) of a flat gradient arena in buckets, overlapped with the backward pass.

    `params[i].grad` must be a view of `flat_grad[offsets[i] : offsets[i] + params[i].numel()]` (FlatArena keeps
    that invariant).  Usage per backward:  sync.begin();  loss.backward();  sync.finish().
    """

    def __init__(self, params: Sequence[torch.Tensor], offsets: Sequence[int], flat_grad: torch.Tensor, group=None,
                 bucket_bytes: int = 4 << 20):
        self.params, self.flat_grad, self.group = list(params), flat_grad, group
        self.buckets = plan_buckets(list(offsets), [p.numel() for p in params], flat_grad.numel(), max(1, bucket_bytes // 4))
        self._bucket_of = {}
        for b, (_, _, members) in enumerate(self.buckets):
            for i in members:
                self._bucket_of[i] = b
        self._pending: List[int] = []
        self._launched: List[bool] = []
        self._works = []
        self._active = False
        self._stream = torch.cuda.Stream(device=flat_grad.device) if flat_grad.is_cuda else None
        self._hooks = [p.register_post_accumulate_grad_hook(self._make_hook(i)) for i, p in enumerate(self.params)
                       if p.requires_grad]

    def _make_hook(self, i):
        def hook(_p):
            if not self._active:
                return
            b = self._bucket_of[i]
            self._pending[b] -= 1
            if self._pending[b] == 0:
                self._launch(b)
        return hook

    def begin(self, expected: Optional[Sequence[bool]] = None):
        """Arm the hooks.  `expected[i]` = parameter i receives a gradient in this backward (default: requires_grad)."""
        exp = [p.requires_grad for p in self.params] if expected is None else list(expected)
        self._pending = [sum(1 for i in members if exp[i]) for (_, _, members) in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._works = []
        self._active = True

    def _launch(self, b):
        if self._launched[b]:
            return
        self._launched[b] = True
        lo, hi, _ = self.buckets[b]
        chunk = self.flat_grad[lo:hi]
        if self._stream is not None:
            self._stream.wait_stream(torch.cuda.current_stream(self.flat_grad.device))
            with torch.cuda.stream(self._stream):
                dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group)
        else:
            self._works.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """Reduce whatever no hook has launched yet (parameters without gradient this pass keep their zeros, which
        still have to be summed for the replicas to stay identical) and join the side stream."""
        self._active = False
        for b in range(len(self.buckets)):
            self._launch(b)
        if self._stream is not None:
            torch.cuda.current_stream(self.flat_grad.device).wait_stream(self._stream)
        for w in self._works:
            w.wait()
        self._works = []

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def sync_count_and_guard(n_valid: torch.Tensor, guard_min: torch.Tensor, guard_max: torch.Tensor, group=None):
    """Make the valid-step count (SUM) and the guard bounds (MIN / MAX) global, in place -- with ONE collective: every
    rank contributes (n_valid, min, max) to an all-gather of 3 doubles and reduces the gathered table locally (three
    separate all-reduces cost three NCCL launch latencies for 20 bytes)."""
    world = dist.get_world_size(group)
    mine = torch.stack((n_valid.reshape(()).to(torch.float64), guard_min.reshape(()).to(torch.float64),
                        guard_max.reshape(()).to(torch.float64)))
    rows = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(rows, mine, group=group)
    table = torch.stack(rows)
    n_valid.copy_(table[:, 0].sum().to(n_valid.dtype).reshape(n_valid.shape))
    guard_min.copy_(table[:, 1].min().to(guard_min.dtype).reshape(guard_min.shape))
    guard_max.copy_(table[:, 2].max().to(guard_max.dtype).reshape(guard_max.shape))

"""Data-parallel gradient synchronisation for the one-process-per-GPU launch (new work: the reference is
single-process).

Pure data parallelism: every rank holds the full parameter store and processes its own shard of samples; the
only exchange is one gradient all-reduce per step.  Because parameters, gradients and optimizer state are flat,
contiguous buffers (params.ParamStore), a "bucket" is simply an element range of the flat gradient buffer -- no
flatten / unflatten copies.  Overlap with backward comes from the order in which the engine finishes gradients:

    head  ->  text tower (ends with the embedding tables, 92 M of the 162 M parameters)  ->  image tower

so the text ranges are all-reduced (on NCCL's stream) while the image tower's backward is still running, and only
the image-tower ranges remain as an exposed tail.  The 1/world_size averaging is folded into the fused Adam kernel
(grad_scale), so no extra pass touches the gradients.  BatchNorm statistics stay per replica, exactly as N
independent runs of the reference would behave (no SyncBN in the reference).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def contiguous_ranges(specs, names, predicate, align=64):
    """Merge the (padded) element ranges of all parameters whose name satisfies ``predicate``."""
    spans = []
    for n in names:
        if predicate(n):
            s = specs[n]
            spans.append([s.offset, s.offset + (s.numel + align - 1) // align * align])
    spans.sort()
    merged = []
    for a, b in spans:
        if merged and merged[-1][1] == a:
            merged[-1][1] = b
        else:
            merged.append([a, b])
    return [tuple(m) for m in merged]


def split_ranges(ranges, max_elems):
    out = []
    for a, b in ranges:
        while b - a > max_elems:
            out.append((a, a + max_elems))
            a += max_elems
        if b > a:
            out.append((a, b))
    return out


class GradSync:
    """All-reduce of the flat gradient buffer in phases that follow the backward's completion order."""

    def __init__(self, store, group=None, bucket_elems: int = 64 * 1024 * 1024, phase_predicates=None):
        self.store = store
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        names = store.names()
        if phase_predicates is None:
            phase_predicates = {"text": lambda n: n.startswith("bert."), "rest": lambda n: not n.startswith("bert.")}
        self.phases = {k: split_ranges(contiguous_ranges(store.specs, names, pred), bucket_elems)
                       for k, pred in phase_predicates.items()}
        self._pending = []

    def covered(self):
        return sorted(r for rs in self.phases.values() for r in rs)

    def ready(self, phase: str):
        """Called by the engine as soon as every gradient of ``phase`` is final: launch its all-reduces."""
        if self.world == 1:
            return
        for a, b in self.phases[phase]:
            self._pending.append(dist.all_reduce(self.store.grad[a:b], op=dist.ReduceOp.SUM, group=self.group,
                                                 async_op=True))

    def finish(self) -> float:
        """Block the current stream on the outstanding all-reduces. Returns the scale that turns the summed
        gradients into the mean (to be folded into the optimizer step)."""
        for w in self._pending:
            w.wait()
        self._pending.clear()
        return 1.0 / self.world

    def broadcast_parameters(self, buffers=()):
        if self.world == 1:
            return
        dist.broadcast(self.store.master, src=0, group=self.group)
        for b in buffers:
            dist.broadcast(b, src=0, group=self.group)
